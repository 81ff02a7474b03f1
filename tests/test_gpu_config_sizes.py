"""Parity at the sizes BASELINE.json names (configs C1..C5), and the two-phase / flat traversals
against the reference-order traversal ON THE DEVICE for every ray of the big configurations.

    C1  scenes/cornel_box.pbrt 900 x 900, 16 spp              whole film against the oracle
    C2  scenes/sphere.pbrt 1600 x 900 (1024 spp in the file)  64 spp, oracle on five crops (centre + corners)
    C3  Cornell box + box-texture.png, 1920 x 1080            64 spp, oracle on five crops
    C4  1,000,008 triangles (plymesh), 1920 x 1080            every camera ray + 2^22 random rays, shipped
                                                              traversal == reference-order traversal (ids, t bits)
    C5  10,000,008 triangles (plymesh), 3840 x 2160           BVH order == oracle, 2^15 rays == oracle,
                                                              every camera ray + 2^22 random rays device == device
plus an adversarial batch (huge / tiny coordinates, coincident surfaces a few ulp apart, slivers).
The oracle cannot trace millions of rays through 10^7 triangles in test time; the reference-order
kernel (k_batch_closest<1>, whose visit counters equal the oracle's) can, in milliseconds.
"""
import importlib
import os
import time

import numpy as np
import pytest

from conftest import SCENES, c3_text, lcg_rays, scene_text

pytestmark = pytest.mark.gpu
NO_HIT = 0xFFFFFFFF


def _relerr(a, b):
    return np.abs(a - b) / (np.abs(b) + 1e-3)


def _crops(W, H, cw, ch):
    return [((W - cw) // 2, (H - ch) // 2), (0, 0), (W - cw, 0), (0, H - ch), (W - cw, H - ch)]


def _compare_crops(r, w, W, H, cw, ch, frac_allowed):
    g, _ = r.render()
    worst = 0.0
    for x0, y0 in _crops(W, H, cw, ch):
        o, _, _ = w.render(rect=(x0, y0, x0 + cw, y0 + ch))
        a, b = g[y0:y0 + ch, x0:x0 + cw], o[y0:y0 + ch, x0:x0 + cw]
        assert np.isfinite(a).all()
        bad = float((_relerr(a, b).max(axis=2) > 1e-3).mean())
        worst = max(worst, bad)
        rmse = float(np.sqrt(np.mean((a - b) ** 2)))
        assert rmse <= 0.02 * float(b.mean()) + 1e-4, (x0, y0, rmse)
    assert worst <= frac_allowed, worst
    return worst


def test_c1_cornell_900x900_16spp_whole_film(cpbrt, oracle, make_scene):
    scene = make_scene("cornel_box.pbrt", spp=16)
    assert scene.resolution == (900, 900)
    r = cpbrt.Renderer(0); r.upload(scene)
    w = oracle.OracleWorld(scene)
    assert np.array_equal(r.build_bvh(), w.bvh_order())
    g, gs = r.render()
    o, _, os_ = w.render()
    bad = float((_relerr(g, o).max(axis=2) > 1e-3).mean())
    print("C1 pixels beyond 1e-3: %.5f%%" % (100 * bad))
    assert bad <= 2e-3, bad
    assert float(np.sqrt(np.mean((g - o) ** 2))) <= 0.01 * float(o.mean())
    assert gs.nan_dropped == os_["nan_dropped"]
    expect = os_["rays_closest"] - os_["rays_unused_last"]
    assert abs(gs.rays_closest - expect) <= 1e-3 * expect and abs(gs.rays_any - os_["rays_any"]) <= 1e-3 * os_["rays_any"]
    # first-hit ids of every sample-0 camera ray at full resolution
    rays = r.camera_rays(0)
    a, b = r.trace_closest(rays), w.trace_closest(rays)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))


def test_c2_sphere_1600x900_crops(cpbrt, oracle, make_scene):
    scene = make_scene("sphere.pbrt", spp=64)
    assert scene.resolution == (1600, 900)
    r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
    w = oracle.OracleWorld(scene)
    print("C2 worst crop fraction beyond 1e-3: %.5f%%" % (100 * _compare_crops(r, w, 1600, 900, 160, 96, 2e-3)))


def test_c3_textured_cornell_1080p_crops(cpbrt, oracle):
    scene = cpbrt.Scene.parse(c3_text().replace('"integer pixelsamples" [1024]', '"integer pixelsamples" [64]'), base_dir=SCENES)
    assert scene.resolution == (1920, 1080) and scene.desc.n_texels > 0
    r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
    w = oracle.OracleWorld(scene)
    print("C3 worst crop fraction beyond 1e-3: %.5f%%" % (100 * _compare_crops(r, w, 1920, 1080, 128, 72, 2e-3)))


def _device_vs_device(r, rays, label):
    """Shipped traversal against the reference-order traversal, both on the device: ids and distance bits."""
    ref, st = r.trace_closest(rays, count=True)          # bvh_closest: the reference's own visiting order
    got = r.trace_closest(rays)
    same = np.array_equal(got["prim"], ref["prim"]) and np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    if not same:
        bad = np.nonzero((got["prim"] != ref["prim"]) | (got["t"].view(np.uint32) != ref["t"].view(np.uint32)))[0]
        raise AssertionError("%s: %d of %d rays differ, first %d: got prim %d t %r, reference order prim %d t %r" % (
            label, len(bad), len(rays), bad[0], got["prim"][bad[0]], got["t"][bad[0]], ref["prim"][bad[0]], ref["t"][bad[0]]))
    return float((ref["prim"] != NO_HIT).mean()), st


def _shadow_device_vs_device(r, rays, label):
    ref, _ = r.trace_any(rays, count=True)
    got = r.trace_any(rays)
    assert np.array_equal(got, ref), "%s: %d shadow rays differ" % (label, int((got != ref).sum()))
    return float(ref.mean())


def _heightfield_exhaustive(cpbrt, oracle, name, oracle_rays):
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    cfg = dict(syn.CONFIGS[name]); cfg["spp"] = 1
    t = {}
    scene = syn.load_heightfield(cpbrt, timing=t, **cfg)
    r = cpbrt.Renderer(0); r.upload(scene)
    t0 = time.time(); order = r.build_bvh(); build_s = time.time() - t0
    print("%s: %d primitives, ply+pbrt written in %.1f s, parsed in %.1f s, BVH built in %.1f s" % (name, scene.desc.n_primitives, t["write_s"], t["parse_s"], build_s))
    w = oracle.OracleWorld(scene)
    assert np.array_equal(order, w.bvh_order()), "BVH primitive order differs from the oracle"
    cam = r.camera_rays(0)
    assert len(cam) == cfg["xres"] * cfg["yres"]
    rnd = lcg_rays(1 << 22, (0, 0, 0), (556, 548, 559), seed=7)
    # the oracle on a subset (its reference-order visit counters must equal the device's) ...
    sub = np.concatenate([cam[:: max(1, len(cam) // (oracle_rays // 2))][: oracle_rays // 2], rnd[: oracle_rays // 2]])
    g, gs = r.trace_closest(sub, count=True)
    o, oc = w.trace_closest(sub, count=True)
    assert np.array_equal(g["prim"], o["prim"]) and np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    assert gs.trace_parent_visits == int(oc[0]) and gs.trace_prim_tests == int(oc[1])
    # ... and every ray device against device
    assert _device_vs_device(r, cam, name + " camera rays")[0] > 0.4
    assert _device_vs_device(r, rnd, name + " random rays")[0] > 0.3
    sh = lcg_rays(1 << 22, (0, 0, 0), (556, 548, 559), seed=13)
    sh[:, 3] = np.float32(1 - 1e-5); sh[:, 4:7] *= np.float32(300)
    assert 0.05 < _shadow_device_vs_device(r, sh, name + " shadow rays") < 0.95
    assert np.array_equal(r.trace_any(sh[: oracle_rays]), w.trace_any(sh[: oracle_rays]))
    r.close()


def test_c4_every_ray_device_vs_device(cpbrt, oracle):
    _heightfield_exhaustive(cpbrt, oracle, "C4", 1 << 15)


def test_c5_ten_million_triangles(cpbrt, oracle):
    """10,000,008 primitives at 3840 x 2160: SAH primitive order (src/aggregate/bvh.rs:24-124) at that size,
    2^15 rays against the oracle, 8.3 M camera rays + 2^22 random rays + 2^22 shadow rays device against device."""
    _heightfield_exhaustive(cpbrt, oracle, "C5", 1 << 15)


# ---- adversarial geometry -------------------------------------------------------------------------
def _soup_text(tris, fov=60):
    P = " ".join(repr(float(v)) for v in np.asarray(tris, np.float32).ravel())
    I = " ".join(str(i) for i in range(3 * len(tris)))
    return ('Camera "perspective" "float fov" [%g]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [64] "integer yresolution" [64]\nIntegrator "path"\nWorldBegin\n'
            'Material "matte"\nShape "trianglemesh" "integer indices" [%s] "point P" [%s]\nWorldEnd\n' % (fov, I, P))


def _adversarial_tris(kind, n, rng):
    if kind == "soup":
        c = rng.random((n, 1, 3)) * 2 - 1
        return c + rng.normal(scale=0.08, size=(n, 3, 3)) + [0, 0, 4]
    if kind == "coincident":
        # stacks of identical quads 1..64 ulp apart along z, plus exactly coincident copies
        out = []
        for k in range(n // 8):
            x, y = rng.random(2) * 1.6 - 0.8
            z = np.float32(3.0 + rng.random())
            for ulps in (0, 0, 1, 2, 5, 17, 64, 3):
                zz = z
                for _ in range(ulps):
                    zz = np.nextafter(zz, np.float32(10))
                s = 0.15
                out.append([[x - s, y - s, zz], [x + s, y - s, zz], [x + s, y + s, zz]])
        return np.array(out, np.float64)
    if kind == "slivers":
        # long needles: one edge 1e-6 .. 1e-3 of the others
        a = rng.random((n, 3)) * 2 - 1 + [0, 0, 4]
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        e = rng.normal(size=(n, 3)); e /= np.linalg.norm(e, axis=1, keepdims=True)
        width = 10.0 ** rng.uniform(-6, -3, size=(n, 1))
        return np.stack([a, a + d * 1.5, a + d * 1.5 + e * width], 1)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["soup", "coincident", "slivers"])
@pytest.mark.parametrize("n", [48, 200, 4096])        # flat search, loop schedule, vote schedule
@pytest.mark.parametrize("scale", [1.0, 1e3, 1e-2, 1e6])        # 1e-2: just beyond the camera's near plane (the clipper drops what is nearer)
def test_adversarial_geometry_all_schedules(cpbrt, oracle, kind, n, scale):
    """The margins of the two-phase and flat searches are relative to the scene's extent; this checks them
    where they could break: coordinates of 1e3..1e6 and 1e-2, coincident surfaces 0..64 ulp apart, needle
    triangles -- against the oracle AND against the reference-order traversal on the device."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(("%s %d" % (kind, n)).encode()))
    tris = (_adversarial_tris(kind, n, rng) * scale).astype(np.float32)
    scene = cpbrt.Scene.parse(_soup_text(tris))
    if scene.desc.n_primitives < 8:
        pytest.skip("the camera clipper kept almost nothing")
    r = cpbrt.Renderer(0); r.upload(scene)
    w = oracle.OracleWorld(scene)
    assert np.array_equal(r.build_bvh(), w.bvh_order())
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    ext = hi - lo
    batches = [r.camera_rays(0), lcg_rays(1 << 16, lo - 0.3 * ext, hi + 0.3 * ext, seed=11)]
    # rays aimed at vertices and edge midpoints of the triangles (where rivals and grazing hits live)
    kept = np.array([[list(scene.desc.triangles[i].p[k]) for k in range(3)] for i in range(min(scene.desc.n_triangles, 1500))], np.float64)
    targets = np.concatenate([kept.reshape(-1, 3), (kept + np.roll(kept, 1, axis=1)).reshape(-1, 3) / 2])
    org = lo + rng.random((len(targets), 3)) * ext - [0, 0, 1.5 * ext[2]]
    aim = np.zeros((len(targets), 8), np.float32)
    aim[:, 0:3] = org; aim[:, 3] = np.finfo(np.float32).max; aim[:, 4:7] = targets - org
    batches.append(aim)
    for i, rays in enumerate(batches):
        _device_vs_device(r, rays, "%s n=%d scale=%g batch %d" % (kind, n, scale, i))
        o = w.trace_closest(rays[:20000])
        g = r.trace_closest(rays[:20000])
        assert np.array_equal(g["prim"], o["prim"]) and np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
        sh = rays.copy(); sh[:, 3] = np.float32(0.999)
        _shadow_device_vs_device(r, sh, "%s n=%d scale=%g shadow batch %d" % (kind, n, scale, i))
        assert np.array_equal(r.trace_any(sh[:20000]), w.trace_any(sh[:20000]))
    r.close()
