"""The synthetic benchmark configurations C4 / C5 as files the reference
itself reads: a height-field mesh written as binary little-endian PLY (`Shape "plymesh"`,
src/geometry/shape/plymesh.rs:14-131: float x y z nx ny nz u v per vertex, `uint8 int` face lists)
under the Cornell box's light, side walls and back wall, seen by the Cornell camera.

    scene = synthetic.load_heightfield(cpbrt, **synthetic.CONFIGS["C4"])   # writes <dir>/c4.pbrt + .ply, loads them

Height field: y = 40 * noise(x, z) over x in [0, 556], z in [0, 559]; noise = a 65 x 65
lattice of 64-bit LCG values (state <- state * 6364136223846793005 + 1442695040888963407, value =
top 24 bits / 2^24, row-major in z then x) interpolated bilinearly in f32; vertices row-major in z;
quad (i, j) gives the triangles (v, v+1, v+nx+2) and (v+nx+2, v+nx+1, v) with v = j*(nx+1)+i."""
import os
import tempfile
import time

import numpy as np

CONFIGS = {
    "C4": dict(nx=1000, nz=500, xres=1920, yres=1080, spp=256),     # 1,000,000 triangles
    "C5": dict(nx=2500, nz=2000, xres=3840, yres=2160, spp=4096),   # 10,000,000 triangles
}


def heightfield_vertices(nx, nz, seed=20260418, height=40.0, xmax=556.0, zmax=559.0):
    """(nz+1)*(nx+1) x 3 float32, every operation in f32 and in a fixed order (a file written on one
    machine must hold the same bits as one written on another)."""
    L = 64
    f32 = np.float32
    m = (L + 1) * (L + 1)
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    with np.errstate(over="ignore"):
        pw = np.cumprod(np.full(m, a, np.uint64))
        geo = np.cumsum(np.concatenate([[np.uint64(1)], pw[:-1]]))
        x = pw * np.uint64(seed) + c * geo
    lat = ((x >> np.uint64(40)).astype(f32) / f32(16777216.0)).reshape(L + 1, L + 1)
    fi = np.arange(nx + 1, dtype=f32) / f32(nx)
    fj = np.arange(nz + 1, dtype=f32) / f32(nz)
    u, w = fi * f32(L), fj * f32(L)
    iu = np.minimum(u.astype(np.int64), L - 1); iw = np.minimum(w.astype(np.int64), L - 1)
    fu = (u - iu.astype(f32))[None, :]; fw = (w - iw.astype(f32))[:, None]
    A = lat[iw[:, None], iu[None, :]]; B = lat[iw[:, None], iu[None, :] + 1]
    C = lat[iw[:, None] + 1, iu[None, :]]; D = lat[iw[:, None] + 1, iu[None, :] + 1]
    n0 = A + (B - A) * fu
    n1 = C + (D - C) * fu
    P = np.empty((nz + 1, nx + 1, 3), f32)
    P[:, :, 0] = (f32(xmax) * fi)[None, :]
    P[:, :, 1] = f32(height) * (n0 + (n1 - n0) * fw)
    P[:, :, 2] = (f32(zmax) * fj)[:, None]
    return P.reshape(-1, 3)


def heightfield_faces(nx, nz):
    j, i = np.meshgrid(np.arange(nz, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    v = (j * (nx + 1) + i).reshape(-1)
    return np.stack([v, v + 1, v + nx + 2, v + nx + 2, v + nx + 1, v], 1).reshape(-1, 3).astype(np.int32)


def write_ply(path, vertices, faces):
    """Binary little-endian PLY with exactly the header lines plymesh.rs expects."""
    nv, nf = len(vertices), len(faces)
    rec = np.zeros((nv, 8), "<f4")
    rec[:, 0:3] = vertices
    rec[:, 4] = 1.0          # normals / uvs are read and ignored by the reference (plymesh.rs:82-131 keeps positions only)
    fr = np.zeros(nf, dtype=[("n", "u1"), ("v", "<i4", 3)])
    fr["n"] = 3; fr["v"] = faces
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % nv +
              "".join("property float %s\n" % p for p in ("x", "y", "z", "nx", "ny", "nz", "u", "v")) +
              "element face %d\nproperty list uint8 int vertex_indices\nend_header\n" % nf)
    tmp = path + ".tmp.%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(header.encode()); f.write(rec.tobytes()); f.write(fr.tobytes())
    os.replace(tmp, path)


def heightfield_scene_text(ply_name, xres, yres, spp, integrator="path", max_depth=5):
    quad = lambda p: 'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [%s]\n' % " ".join("%g" % v for v in p)
    return (
        'LookAt 278 273 -800 278 273 0 0 1 0\n'
        'Camera "perspective" "float fov" [37.5]\n'
        'Sampler "halton" "integer pixelsamples" [%d]\n'
        'Film "image" "string filename" ["heightfield.png"] "integer xresolution" [%d] "integer yresolution" [%d]\n'
        'Integrator "%s" "integer maxdepth" [%d]\n'
        'WorldBegin\n'
        'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [17 12 4]\n%sAttributeEnd\n'
        'Material "matte" "rgb Kd" [0.582553 0.043657 0.044207]\n%s'
        'Material "matte" "rgb Kd" [0.100772 0.376082 0.078298]\n%s'
        'Material "matte" "rgb Kd" [0.886975 0.696752 0.667831]\n%s'
        'Material "matte" "rgb Kd" [0.5 0.5 0.5]\n'
        'Shape "plymesh" "string filename" ["%s"]\n'
        'WorldEnd\n' % (
            spp, xres, yres, integrator, max_depth,
            quad([343, 548.7, 227, 343, 548.7, 332, 213, 548.7, 332, 213, 548.7, 227]),
            quad([552.8, 0, 0, 549.6, 0, 559.2, 556, 548.8, 559.2, 556, 548.8, 0]),
            quad([0, 0, 559.2, 0, 0, 0, 0, 548.8, 0, 0, 548.8, 559.2]),
            quad([549.6, 0, 559.2, 0, 0, 559.2, 0, 548.8, 559.2, 556, 548.8, 559.2]),
            ply_name))


def write_heightfield_files(directory, nx, nz, xres, yres, spp, seed=20260418, integrator="path", max_depth=5):
    """Writes <directory>/heightfield_<nx>x<nz>_<seed>.ply (once) and a .pbrt next to it; returns the .pbrt path."""
    os.makedirs(directory, exist_ok=True)
    ply = "heightfield_%dx%d_%d.ply" % (nx, nz, seed)
    if not os.path.exists(os.path.join(directory, ply)):
        write_ply(os.path.join(directory, ply), heightfield_vertices(nx, nz, seed), heightfield_faces(nx, nz))
    pbrt = os.path.join(directory, "heightfield_%dx%d_%d_%dx%d_%dspp_%s%d.pbrt" % (nx, nz, seed, xres, yres, spp, integrator, max_depth))
    text = heightfield_scene_text(ply, xres, yres, spp, integrator, max_depth)
    try:
        with open(pbrt) as f:
            if f.read() == text:
                return pbrt          # several ranks of one box call this at once: never truncate a file a neighbour is reading
    except OSError:
        pass
    tmp = pbrt + ".tmp.%d" % os.getpid()
    with open(tmp, "w") as f:
        f.write(text)
    os.replace(tmp, pbrt)
    return pbrt


def default_dir():
    return os.environ.get("CPBRT_SYNTHETIC_DIR") or os.path.join(tempfile.gettempdir(), "cpbrt_synthetic")


def load_heightfield(cpbrt, nx, nz, xres, yres, spp, seed=20260418, integrator="path", max_depth=5, directory=None, timing=None):
    """The configuration as a parsed Scene, through the same `Scene.load(path)` a user's .pbrt takes.
    `timing` (a dict) receives the seconds spent writing the files and parsing them."""
    t0 = time.time()
    pbrt = write_heightfield_files(directory or default_dir(), nx, nz, xres, yres, spp, seed, integrator, max_depth)
    t1 = time.time()
    scene = cpbrt.Scene.load(pbrt)
    if timing is not None:
        timing.update(write_s=t1 - t0, parse_s=time.time() - t1, pbrt=pbrt)
    return scene
