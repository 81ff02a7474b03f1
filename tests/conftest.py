import importlib
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def cpbrt():
    """The product package (directory name has a hyphen, hence importlib)."""
    _ensure_built()
    return importlib.import_module("curry-pbrt_b200")


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    import oracle_api
    return oracle_api


SCENES = os.path.join(ROOT, "scenes")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def scene_text(name, xres=None, yres=None, spp=None, integrator=None):
    """Scene file text with the Film / Sampler / Integrator lines rewritten (the reference takes
    these from the file only)."""
    text = open(os.path.join(SCENES, name)).read()
    if xres is not None:
        text = re.sub(r'"integer xresolution" \[\s*\d+\s*\]', '"integer xresolution" [%d]' % xres, text)
    if yres is not None:
        text = re.sub(r'"integer yresolution" \[\s*\d+\s*\]', '"integer yresolution" [%d]' % yres, text)
    if spp is not None:
        text = re.sub(r'"integer pixelsamples" \[\s*\d+\s*\]', '"integer pixelsamples" [%d]' % spp, text)
    if integrator is not None:
        text = re.sub(r'Integrator "[a-z]+"\s*"integer maxdepth" \[\s*\d+\s*\]', 'Integrator "%s" "integer maxdepth" [%d]' % integrator, text)
    return text


def c3_text(scene_text=None):
    """BASELINE config C3: the Cornell box with box-texture.png as the Kd of its white surfaces, 1920x1080, 1024 spp."""
    text = globals()["scene_text"]("cornel_box.pbrt", xres=1920, yres=1080, spp=1024)
    text = text.replace('WorldBegin\n', 'WorldBegin\nTexture "boxtex" "spectrum" "imagemap" "string filename" "box-texture.png"\n', 1)
    parts = text.split('Material "matte"\n"spectrum Kd"')
    head = 'Material "matte"\n"spectrum Kd"'.join(parts[:4])
    rest = parts[4]
    rest = rest[rest.index(']') + 1:]
    return head + 'Material "matte" "texture Kd" "boxtex"' + rest


@pytest.fixture(scope="session")
def make_scene(cpbrt):
    def _make(name, **kw):
        return cpbrt.Scene.parse(scene_text(name, **kw), base_dir=SCENES)
    return _make


def lcg_rays(n, lo, hi, seed=7):
    """Fixed ray batch: origins uniform in [lo, hi], directions uniform on the
    sphere, driven by a 64-bit LCG so the batch is identical everywhere."""
    m = n * 5
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    with np.errstate(over="ignore"):
        # x_k = a^k x_0 + c (1 + a + ... + a^(k-1))  (mod 2^64), k = 1..m
        pw = np.cumprod(np.full(m, a, np.uint64))                    # a^1 .. a^m
        geo = np.cumsum(np.concatenate([[np.uint64(1)], pw[:-1]]))   # 1 + a + ... + a^(k-1)
        x = pw * np.uint64(seed) + c * geo
    vals = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    v = vals.reshape(n, 5)
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64)
    o = lo + v[:, :3] * (hi - lo)
    z = 1.0 - 2.0 * v[:, 3]
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = 2.0 * np.pi * v[:, 4]
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o; rays[:, 3] = np.finfo(np.float32).max; rays[:, 4:7] = d
    return rays


def write_exr(path, img, compression, pixel_type, origin=(0, 0)):
    """Minimal OpenEXR scanline writer (test fixture): img is (H, W, 3) float32 in R, G, B order;
    compression 0 NONE / 1 RLE / 2 ZIPS / 3 ZIP; pixel_type 0 UINT / 1 HALF / 2 FLOAT."""
    import struct
    import zlib
    H, W, _ = img.shape
    x0, y0 = origin

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    chlist = b"".join(n + b"\0" + struct.pack("<iBxxxii", pixel_type, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<4i", x0, y0, x0 + W - 1, y0 + H - 1)
    header = struct.pack("<II", 20000630, 2)
    header += attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([compression]))
    header += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box)
    header += attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    header += b"\0"
    dt = {0: np.uint32, 1: np.float16, 2: np.float32}[pixel_type]
    lines = 16 if compression == 3 else 1

    def rle(b):
        out, i = bytearray(), 0
        while i < len(b):
            j = i
            while j + 1 < len(b) and b[j + 1] == b[i] and j - i < 126:
                j += 1
            if j - i >= 2:
                out += struct.pack("b", j - i) + bytes([b[i]])
                i = j + 1
            else:
                k = i
                while k < len(b) and k - i < 127 and not (k + 2 < len(b) and b[k] == b[k + 1] == b[k + 2]):
                    k += 1
                out += struct.pack("b", -(k - i)) + bytes(b[i:k])
                i = k
        return bytes(out)

    blocks = []
    for r0 in range(0, H, lines):
        raw = b"".join(img[r, :, c].astype(dt).tobytes() for r in range(r0, min(H, r0 + lines)) for c in (2, 1, 0))
        data = raw
        if compression:
            a = np.frombuffer(raw, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
            p = t.copy()
            p[1:] = (t[1:] - t[:-1] + 128) & 0xFF
            packed = rle(bytes(p.astype(np.uint8))) if compression == 1 else zlib.compress(bytes(p.astype(np.uint8)))
            if len(packed) < len(raw):
                data = packed
        blocks.append(struct.pack("<ii", y0 + r0, len(data)) + data)
    table_at = len(header)
    offsets, pos = [], table_at + 8 * len(blocks)
    for b in blocks:
        offsets.append(pos)
        pos += len(b)
    with open(path, "wb") as f:
        f.write(header + struct.pack("<%dQ" % len(blocks), *offsets) + b"".join(blocks))


def specular_furnace_scene_text(material, spp=256, res=24):
    """White furnace for the delta lobes: a mirror sphere (Kr = 1) or a glass sphere with Kr = Kt = 1 (the
    reference's default Kr is 0.5) returns exactly the environment's radiance -- Fresnel reflectance plus
    transmittance is 1 at every interface, so this checks fresnel_dielectric, refract, the luminance-weighted
    choice between the two delta lobes and its pmf, and Russian roulette; only paths cut at maxdepth are lost."""
    mat = 'Material "mirror"' if material == "mirror" else 'Material "glass" "rgb Kr" [1 1 1]'
    return ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\nIntegrator "path" "integer maxdepth" [12]\nWorldBegin\n'
            'LightSource "infinite"\n%s\nTranslate 0 0 5\nShape "sphere" "float radius" [1.2]\nWorldEnd\n' % (spp, res, res, mat))


def furnace_scene_text(integrator, kd, spp=256, res=24):
    """White furnace: a convex matte sphere inside a constant environment of radiance 1.  No path can
    bounce twice on a convex body, so the exitant radiance is exactly Kd wherever the sphere covers the
    pixel and exactly 1 elsewhere, whatever the integrator, the sampling strategy or the MIS weights."""
    return ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\nIntegrator "%s" "integer maxdepth" [5]\nWorldBegin\n'
            'LightSource "infinite"\nMaterial "matte" "rgb Kd" [%g %g %g]\nTranslate 0 0 5\nShape "sphere" "float radius" [1.2]\nWorldEnd\n'
            % (spp, res, res, integrator, kd, kd * 0.5, kd * 0.25))


def point_light_quad_scene_text(integrator):
    """A matte square facing the camera at distance 5 with a point light at the eye: the radiance through a
    pixel at angle a off the axis is Kd / pi * I * cos(a)^3 / 25 (inverse square law at 5 / cos a, cosine
    at the surface) -- closed form for the delta-light branch of uniform_sample_one_light."""
    return ('Camera "perspective" "float fov" [20]\nSampler "halton" "integer pixelsamples" [8]\n'
            'Film "image" "integer xresolution" [33] "integer yresolution" [33]\nIntegrator "%s" "integer maxdepth" [1]\nWorldBegin\n'
            'LightSource "point" "rgb I" [10 5 2.5] "point from" [0 0 0]\nMaterial "matte" "rgb Kd" [0.5 0.5 0.5]\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-0.5 -0.5 5  0.5 -0.5 5  0.5 0.5 5  -0.5 0.5 5]\nWorldEnd\n' % integrator)


def point_light_quad_expected():
    """(mask of pixels well inside the square, expected red channel) for point_light_quad_scene_text."""
    ys, xs = np.mgrid[0:33, 0:33]
    t = np.tan(np.radians(10))
    sx, sy = (-1 + xs / 33 * 2) * t, (1 - ys / 33 * 2) * t
    cos_a = 1 / np.sqrt(sx * sx + sy * sy + 1)
    inside = (np.abs(sx * 5) < 0.45) & (np.abs(sy * 5) < 0.45)
    return inside, 0.5 / np.pi * 10 * cos_a ** 3 / 25


def area_light_quad_scene_text(integrator, spp=64):
    """The same matte square lit by a small diffuse emitter (0.2 x 0.2, two triangle lights) on the axis
    behind the eye, 8 away: irradiance L * A * cos^2 / r^2 (small-source limit, error ~ (0.1 / 8)^2) -- closed
    form for AreaLight::sample_li, the area -> solid-angle pdf, one-light selection and the MIS weights."""
    return ('Camera "perspective" "float fov" [20]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [33] "integer yresolution" [33]\nIntegrator "%s" "integer maxdepth" [1]\nWorldBegin\n'
            'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [50 25 12.5]\nMaterial "matte"\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-0.1 -0.1 -3  0.1 -0.1 -3  0.1 0.1 -3  -0.1 0.1 -3]\nAttributeEnd\n'
            'Material "matte" "rgb Kd" [0.5 0.5 0.5]\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-0.5 -0.5 5  0.5 -0.5 5  0.5 0.5 5  -0.5 0.5 5]\nWorldEnd\n'
            % (spp, integrator))


def area_light_quad_expected():
    ys, xs = np.mgrid[0:33, 0:33]
    t = np.tan(np.radians(10))
    x, y = (-1 + xs / 33 * 2) * t * 5, (1 - ys / 33 * 2) * t * 5
    r2 = x * x + y * y + 64.0
    inside = (np.abs(x) < 0.42) & (np.abs(y) < 0.42)
    return inside, 0.5 / np.pi * 50 * 0.04 * (64.0 / r2) / r2


def distant_light_quad_scene_text(to):
    """The matte square under a distant light of radiance (3, 1.5, 0.75) travelling along `to`: Kd / pi * L * cos."""
    return ('Camera "perspective" "float fov" [20]\nSampler "halton" "integer pixelsamples" [4]\n'
            'Film "image" "integer xresolution" [33] "integer yresolution" [33]\nIntegrator "path" "integer maxdepth" [1]\nWorldBegin\n'
            'LightSource "distant" "point from" [0 0 0] "point to" [%s] "rgb L" [3 1.5 0.75]\nMaterial "matte" "rgb Kd" [0.5 0.5 0.5]\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-0.5 -0.5 5  0.5 -0.5 5  0.5 0.5 5  -0.5 0.5 5]\nWorldEnd\n' % to)


DISTANT_CASES = [("0 0 1", 1.0), ("1 0 1", 0.5 ** 0.5), ("0 -1 1", 0.5 ** 0.5), ("1 1 1", 3 ** -0.5)]


def sphere_light_quad_scene_text(integrator, spp=64):
    """The matte square under a spherical diffuse emitter (radius 0.5, on the axis 8 away): a uniformly emitting
    sphere gives the irradiance pi * L * (R / d)^2 * cos exactly, so the radiance is Kd * L * R^2 / d^2 * cos --
    closed form for Sphere::sample_by_point (cone sampling), by_point_pdf and TransformShape around a light."""
    return ('Camera "perspective" "float fov" [20]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [33] "integer yresolution" [33]\nIntegrator "%s" "integer maxdepth" [1]\nWorldBegin\n'
            'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [50 25 12.5]\nTranslate 0 0 -3\nShape "sphere" "float radius" [0.5]\nAttributeEnd\n'
            'Material "matte" "rgb Kd" [0.5 0.5 0.5]\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-0.5 -0.5 5  0.5 -0.5 5  0.5 0.5 5  -0.5 0.5 5]\nWorldEnd\n'
            % (spp, integrator))


def sphere_light_quad_expected():
    ys, xs = np.mgrid[0:33, 0:33]
    t = np.tan(np.radians(10))
    x, y = (-1 + xs / 33 * 2) * t * 5, (1 - ys / 33 * 2) * t * 5
    r2 = x * x + y * y + 64.0
    inside = (np.abs(x) < 0.42) & (np.abs(y) < 0.42)
    return inside, 0.5 * 50 * 0.25 / r2 * (8 / np.sqrt(r2))
