"""Pins the CPU oracle: the reference's own unit tests, values derived from its sources, an
independent brute-force aggregate, committed golden vectors, and the product's host-side BVH
builder (same exact SAH, different implementation)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import DISTANT_CASES, SCENES, area_light_quad_expected, distant_light_quad_scene_text, area_light_quad_scene_text, furnace_scene_text, lcg_rays, point_light_quad_expected, point_light_quad_scene_text, scene_text, specular_furnace_scene_text, sphere_light_quad_expected, sphere_light_quad_scene_text

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_halton_works_kat(cpbrt, oracle):
    """src/sampler/halton.rs:232-249: HaltonSampler::new(16, (9, 4)), four pixels."""
    s = cpbrt.Scene.parse('Camera "perspective"\nSampler "halton" "integer pixelsamples" [16]\n'
                          'Film "image" "integer xresolution" [9] "integer yresolution" [4]\nIntegrator "path"\nWorldBegin\nWorldEnd\n')
    d = s.desc
    assert list(d.sampler.scale) == [16, 9] and list(d.sampler.exp) == [4, 2]
    expect = {(3, 3): 28, (5, 3): 10, (4, 3): 82, (8, 0): 81}
    for (x, y), index in expect.items():
        got = oracle.lib.orc_halton_pixel_index(C.cast(C.pointer(d.sampler), C.c_void_p), x, y)
        assert got == index
        rx = int(np.float32(oracle.lib.orc_radical_inverse(got, 0)) * np.float32(16))
        ry = int(np.float32(oracle.lib.orc_radical_inverse(got, 1)) * np.float32(9))
        assert (rx, ry) == (x, y)


def test_reference_kats_through_the_second_reading(cpbrt):
    """The reference's two unit tests (halton.rs:232-249, perspective.rs:88-105) against oracle/front_numpy.py, which
    builds the sampler from the resolution by itself (no scene description, no table: dimensions 0 and 1 are unscrambled)."""
    front = _load_oracle_py("front_numpy")
    smp = front.HaltonSampler(16, (9, 4), None)
    assert (smp.scale, smp.exp) == ([16, 9], [4, 2])
    for (x, y), index in {(3, 3): 28, (5, 3): 10, (4, 3): 82, (8, 0): 81}.items():
        smp.set_pixel(x, y)
        assert smp.index == index
        rx = int(np.trunc(front.radical_inverse(smp.index, 0) * front.usize_to_f32(smp.scale[0])))
        ry = int(np.trunc(front.radical_inverse(smp.index, 1) * front.usize_to_f32(smp.scale[1])))
        assert (rx, ry) == (x, y)
    s = cpbrt.Scene.parse('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
                          'Film "image" "integer xresolution" [1024] "integer yresolution" [768]\nIntegrator "path"\nWorldBegin\nWorldEnd\n')
    cam = front.PerspectiveCamera(list(s.desc.camera.raster_to_camera))
    o, d = cam.generate_ray(512.0, 384.0)
    assert [float(v) for v in o + d] == [0.0, 0.0, 0.0, 0.0, 0.0, 1.0]


def test_halton_example_points(cpbrt, oracle):
    """examples/halton_sampler.rs prints get_2d() of samples 0.. for a 1024x768 sampler starting
    at index 0; dims 0,1 involve no permutation."""
    s = cpbrt.Scene.parse('Camera "perspective"\nSampler "halton" "integer pixelsamples" [128]\n'
                          'Film "image" "integer xresolution" [1024] "integer yresolution" [768]\nIntegrator "path"\nWorldBegin\nWorldEnd\n')
    w = oracle.OracleWorld(s)
    # pixel (0,0) has base index 0
    k = np.arange(4, dtype=np.uint32)
    z = np.zeros(4, np.uint32)
    idx, x = w.halton(z, z, k, np.zeros(4, np.uint32))
    _, y = w.halton(z, z, k, np.ones(4, np.uint32))
    assert list(idx) == [i * 1024 * 2187 for i in range(4)]
    np.testing.assert_array_equal(x, np.array([0, 0.8166504, 0.4083252, 0.5218506], np.float32))
    np.testing.assert_array_equal(y, np.array([0, 0.6438043, 0.85962504, 0.21460143], np.float32))


def test_radical_inverse_against_python(cpbrt, oracle):
    def rad_inv(x, base):
        acc, cnt = 0, 0
        while x:
            acc = acc * base + x % base; x //= base; cnt += 1
        return np.float32(np.float32(1) / np.float32(base ** cnt)) * np.float32(acc)
    rng = np.random.RandomState(0)
    for x in list(rng.randint(0, 2 ** 31, 200)) + [0, 1, 2 ** 36 + 12345]:
        for bi, base in ((0, 2), (1, 3)):
            assert np.float32(oracle.lib.orc_radical_inverse(int(x), bi)) == rad_inv(int(x), base)


def test_scrambled_radical_inverse_against_exact_arithmetic(cpbrt, oracle, make_scene):
    """Dimensions >= 2 (halton.rs:56-64): digits of the index in base PRIMS[dim], each replaced through the
    dimension's block of the permutation table, plus the infinite tail of permuted zero digits.  Recomputed
    here with Python integers / Fractions (no rounding at all) from the index the oracle reports; the
    oracle's f32 value must be the exact one to within the rounding of its three f32 operations."""
    from fractions import Fraction
    perm = np.zeros(cpbrt.abi.PERM_TABLE_LEN, np.uint16)
    cpbrt.lib.cpbrt_make_permutation(1, perm.ctypes.data_as(C.POINTER(C.c_uint16)))
    primes = [p for p in range(2, 8000) if all(p % q for q in range(2, int(p ** 0.5) + 1))][:1000]
    offsets = np.concatenate([[0], np.cumsum(primes)])
    w = oracle.OracleWorld(make_scene("cornel_box.pbrt", xres=64, yres=48, spp=4))
    rng = np.random.RandomState(3)
    n = 3000
    px = rng.randint(0, 64, n).astype(np.uint32); py = rng.randint(0, 48, n).astype(np.uint32)
    sample = rng.randint(0, 1 << 22, n).astype(np.uint32)
    dim = rng.randint(2, 1000, n).astype(np.uint32)
    dim[:60] = np.arange(2, 62)
    idx, val = w.halton(px, py, sample, dim)
    assert idx.max() > 2 ** 32
    for i in range(n):
        b, block = primes[dim[i]], perm[offsets[dim[i]]:offsets[dim[i]] + primes[dim[i]]]
        x, rev, digits = int(idx[i]), 0, 0
        while x:
            rev = rev * b + int(block[x % b]); x //= b; digits += 1
        exact = (Fraction(rev) + Fraction(int(block[0]), b - 1)) / Fraction(b) ** digits
        assert abs(Fraction(float(val[i])) - exact) <= exact * Fraction(4, 2 ** 23) + Fraction(1, 2 ** 40), (i, dim[i], float(val[i]), float(exact))


@pytest.mark.parametrize("group_max", [2, 25, 128, 1024, 4096, 65536])
def test_halton_digit_groups_equal_the_digit_loop(cpbrt, group_max):
    """The device takes several digits of scrambled_radical_inverse per round (one division by base^g and one
    look-up in a table of permuted-and-reversed g-digit strings, csrc/device/halton.cuh).  The tables are built on
    the host (csrc/host/halton_tables.cpp) and the same evaluation order runs on the host through
    cpbrt_host_halton_grouped: its integer accumulator and digit count must equal the reference's one-digit loop
    (halton.rs:56-64), recomputed here with Python integers, for every group size -- indices up to 2^40 (the 64-bit
    prefix), indices just below and above powers of the base (partial last groups of every length), dimension 2
    (base 5: 4-5 digits per group) to 999."""
    perm = np.zeros(cpbrt.abi.PERM_TABLE_LEN, np.uint16)
    cpbrt.lib.cpbrt_make_permutation(7, perm.ctypes.data_as(C.POINTER(C.c_uint16)))
    primes = [p for p in range(2, 8000) if all(p % q for q in range(2, int(p ** 0.5) + 1))][:1000]
    offsets = np.concatenate([[0], np.cumsum(primes)])
    rng = np.random.default_rng(group_max)
    n = 6000
    dim = rng.integers(2, 1000, n).astype(np.uint32)
    dim[:400] = np.arange(400) % 40 + 2                       # the small bases, where groups are longest
    index = rng.integers(0, 1 << 40, n, dtype=np.uint64)
    index[:1500] = rng.integers(0, 1 << 31, 1500, dtype=np.uint64)      # what a 900 x 900 x 1024-spp render draws
    for k in range(1500, 2500):                               # b^j - 1, b^j, b^j + 1
        b = primes[dim[k]]
        j = int(rng.integers(1, 14))
        index[k] = max(0, min(b ** j + int(rng.integers(-1, 2)), (1 << 63) - 1))
    index[2500:2520] = np.arange(20)                          # tiny indices, zero included
    acc = np.zeros(n, np.uint64); digits = np.zeros(n, np.uint32)
    rc = cpbrt.lib.cpbrt_host_halton_grouped(perm.ctypes.data, group_max, index.ctypes.data, dim.ctypes.data, n, acc.ctypes.data, digits.ctypes.data)
    assert rc == 0
    for i in range(n):
        b, block = primes[dim[i]], perm[offsets[dim[i]]:offsets[dim[i]] + primes[dim[i]]]
        x, rev, cnt = int(index[i]), 0, 0
        while x:
            rev = (rev * b + int(block[x % b])) % (1 << 64); x //= b; cnt += 1
        assert (int(acc[i]), int(digits[i])) == (rev, cnt), (i, int(index[i]), int(dim[i]))


def test_permutation_table_is_blockwise_permutation(cpbrt):
    perm = np.zeros(cpbrt.abi.PERM_TABLE_LEN, np.uint16)
    cpbrt.lib.cpbrt_make_permutation(1, perm.ctypes.data_as(C.POINTER(C.c_uint16)))
    primes = [p for p in range(2, 8000) if all(p % q for q in range(2, int(p ** 0.5) + 1))][:1000]
    assert sum(primes) == cpbrt.abi.PERM_TABLE_LEN
    off = 0
    for p in primes[:200]:
        assert sorted(perm[off:off + p]) == list(range(p))
        off += p
    perm2 = np.zeros_like(perm)
    cpbrt.lib.cpbrt_make_permutation(2, perm2.ctypes.data_as(C.POINTER(C.c_uint16)))
    assert not np.array_equal(perm, perm2)


def _scene_bbox_rays(n, seed=7):
    return lcg_rays(n, (-50, -50, -50), (610, 600, 610), seed)


def test_bvh_closest_equals_brute_force(make_scene, oracle):
    """Closest hit through the BVH == PlainAggregate-style linear scan (plain.rs:17-36), except
    for exact-distance ties where the two visit orders legitimately pick different primitives."""
    w = oracle.OracleWorld(make_scene("cornel_box.pbrt"))
    rays = _scene_bbox_rays(20000)
    a, b = w.trace_closest(rays), w.trace_closest_plain(rays)
    hit_a, hit_b = a["prim"] != 0xFFFFFFFF, b["prim"] != 0xFFFFFFFF
    assert np.array_equal(hit_a, hit_b)
    assert hit_a.mean() > 0.5
    np.testing.assert_array_equal(a["t"][hit_a], b["t"][hit_a])
    differ = a["prim"] != b["prim"]
    assert differ.mean() < 1e-3          # only ties


@pytest.mark.parametrize("xres,yres", [(900, 900), (1600, 900), (300, 700)])
def test_camera_rays_against_closed_form(make_scene, oracle, xres, yres):
    """PerspectiveCamera::new builds raster_to_camera as the inverse of perspective * translate * scale * scale
    (perspective.rs:10-45) and the scene's LookAt supplies camera_to_world.  Written out by hand in float64:
    d = normalize(right * sx * tan(fov/2) + up * sy * tan(fov/2) + dir) with the screen window [-a, a] x [-1, 1]
    (or [-1, 1] x [-1/a, 1/a] for a < 1), sx = lerp over x / W, sy = top - y / H * height; origin = the eye."""
    w = oracle.OracleWorld(make_scene("cornel_box.pbrt", xres=xres, yres=yres, spp=1))
    a = xres / yres
    half_w, half_h = (a, 1.0) if a > 1 else (1.0, 1.0 / a)
    tan_half = np.tan(np.radians(37.5) / 2)
    rng = np.random.RandomState(9)
    for fx, fy in [(0.0, 0.0), (xres, yres), (xres / 2, yres / 2)] + [tuple(p) for p in rng.rand(40, 2) * [xres, yres]]:
        sx = -half_w + fx / xres * 2 * half_w
        sy = half_h - fy / yres * 2 * half_h
        d = np.array([sx * tan_half, sy * tan_half, 1.0])       # right = +x, up = +y, dir = +z for this LookAt
        d /= np.linalg.norm(d)
        ray = w.camera_ray_at(float(fx), float(fy))
        np.testing.assert_allclose(ray[0:3], [278.0, 273.0, -800.0], atol=1e-3)
        np.testing.assert_allclose(ray[4:7], d, atol=2e-5)


@pytest.mark.parametrize("integrator", ["path", "directlighting"])
def test_white_furnace(cpbrt, oracle, integrator):
    """An analytic answer for the whole estimator (BSDF, cosine sampling, environment sampling and its
    pdf, MIS weights, Russian roulette, film mean): see furnace_scene_text."""
    kd = 0.8
    w = oracle.OracleWorld(cpbrt.Scene.parse(furnace_scene_text(integrator, kd)))
    film, _, st = w.render()
    assert st["nan_dropped"] == 0
    np.testing.assert_allclose(film[10:14, 10:14].mean(axis=(0, 1)), [kd, kd * 0.5, kd * 0.25], rtol=2e-3)
    np.testing.assert_allclose(film[0, 0], [1, 1, 1], rtol=1e-6)


@pytest.mark.parametrize("integrator", ["path", "directlighting"])
def test_point_light_inverse_square_closed_form(cpbrt, oracle, integrator):
    w = oracle.OracleWorld(cpbrt.Scene.parse(point_light_quad_scene_text(integrator)))
    film, _, _ = w.render()
    inside, expect = point_light_quad_expected()
    # the film point of a sample drifts by the accumulated jitter (render.rs:26-31), a fraction of a pixel here
    np.testing.assert_allclose(film[..., 0][inside], expect[inside], rtol=3e-3)
    np.testing.assert_allclose(film[..., 1][inside], expect[inside] / 2, rtol=3e-3)


@pytest.mark.parametrize("integrator", ["path", "directlighting"])
def test_area_light_small_source_closed_form(cpbrt, oracle, integrator):
    w = oracle.OracleWorld(cpbrt.Scene.parse(area_light_quad_scene_text(integrator)))
    film, _, _ = w.render()
    inside, expect = area_light_quad_expected()
    np.testing.assert_allclose(film[..., 0][inside], expect[inside], rtol=3e-3)
    np.testing.assert_allclose(film[..., 2][inside], expect[inside] / 4, rtol=3e-3)


@pytest.mark.parametrize("material", ["mirror", "glass"])
def test_specular_white_furnace(cpbrt, oracle, material):
    w = oracle.OracleWorld(cpbrt.Scene.parse(specular_furnace_scene_text(material)))
    film, _, st = w.render()
    assert st["nan_dropped"] == 0
    np.testing.assert_allclose(film[10:14, 10:14].mean(axis=(0, 1)), [1, 1, 1], rtol=2e-3)
    np.testing.assert_allclose(film[7:17, 7:17].mean(axis=(0, 1)), [1, 1, 1], rtol=3e-3)


@pytest.mark.parametrize("to,cos", DISTANT_CASES)
def test_distant_light_closed_form(cpbrt, oracle, to, cos):
    film, _, _ = oracle.OracleWorld(cpbrt.Scene.parse(distant_light_quad_scene_text(to))).render()
    np.testing.assert_allclose(film[12:21, 12:21].reshape(-1, 3), np.tile(0.5 / np.pi * cos * np.array([3, 1.5, 0.75]), (81, 1)), rtol=2e-6)


@pytest.mark.parametrize("integrator", ["path", "directlighting"])
def test_sphere_light_closed_form(cpbrt, oracle, integrator):
    film, _, _ = oracle.OracleWorld(cpbrt.Scene.parse(sphere_light_quad_scene_text(integrator))).render()
    inside, expect = sphere_light_quad_expected()
    ratio = film[..., 0][inside] / expect[inside]
    assert abs(ratio.mean() - 1) < 3e-3 and np.abs(ratio - 1).max() < 3e-2, (ratio.mean(), ratio.min(), ratio.max())
    np.testing.assert_allclose(film[..., 2][inside], film[..., 0][inside] / 4, rtol=1e-5)


def test_closest_hits_against_float64_geometry(cpbrt, make_scene, oracle):
    """The oracle's closest hits (watertight f32 triangle test, f64-quadratic sphere under a transform,
    BVH with the reference's visiting rules) against a formulation that shares nothing with it:
    Moeller-Trumbore triangles and an analytic sphere, all primitives for every ray, in float64.
    Rays whose answer depends on rounding (two hits within 1e-4 of each other, a hit within 1e-5
    of a triangle edge, grazing incidence) are set aside; every other ray must name the same
    primitive at the same distance."""
    scene = make_scene("cornel_box.pbrt", xres=32, yres=32, spp=1)
    d = scene.desc
    w = oracle.OracleWorld(scene)
    rays = lcg_rays(20000, (0, 0, -100), (556, 548, 559), seed=23)
    o, dr = rays[:, 0:3].astype(np.float64), rays[:, 4:7].astype(np.float64)
    n = len(rays)
    T = np.full((n, int(d.n_primitives)), np.inf)
    edge_margin = np.full((n, int(d.n_primitives)), np.inf)
    for k in range(int(d.n_primitives)):
        sh = d.primitives[k].shape
        oo, dd = o, dr
        x = sh.xform
        chain = []
        while x >= 0:
            chain.append(x); x = d.transforms[x].outer
        for x in reversed(chain):            # world -> object through the inverse matrices, outermost first
            mi = np.array(list(d.transforms[x].m_inv), np.float64).reshape(4, 4)
            oo = oo @ mi[:3, :3].T + mi[:3, 3]; dd = dd @ mi[:3, :3].T
        if sh.kind == cpbrt.abi.SHAPE_TRIANGLE:
            tri = np.array([list(v) for v in d.triangles[sh.data].p], np.float64)
            e1, e2 = tri[1] - tri[0], tri[2] - tri[0]
            pv = np.cross(dd, e2); det = pv @ e1
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / det
                tv = oo - tri[0]; u = (tv * pv).sum(1) * inv
                qv = np.cross(tv, e1); v = (dd * qv).sum(1) * inv
                t = qv @ e2 * inv
            ok = (np.abs(det) > 1e-9) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 1e-6)
            T[ok, k] = t[ok]
            cosang = np.abs(det) / (np.linalg.norm(dd, axis=1) * np.linalg.norm(np.cross(e1, e2)))
            m = np.minimum(np.minimum(np.abs(u), np.abs(v)), np.abs(1 - u - v))
            m = np.where(cosang < 1e-3, 0.0, m)          # grazing rays count as ambiguous
            edge_margin[:, k] = np.where(np.isfinite(m), m, 0.0)
        else:
            r = float(d.sphere_radius[sh.data])
            a = (dd * dd).sum(1); b = 2 * (dd * oo).sum(1); c = (oo * oo).sum(1) - r * r
            disc = b * b - 4 * a * c
            sq = np.sqrt(np.maximum(disc, 0))
            t0, t1 = (-b - sq) / (2 * a), (-b + sq) / (2 * a)
            t = np.where(t0 > 1e-6, t0, t1)
            ok = (disc > 0) & (t > 1e-6)
            T[ok, k] = t[ok]
            edge_margin[:, k] = np.where(disc > 0, sq / np.maximum(np.abs(b), 1e-30), 0.0)   # tangent rays are ambiguous
    best = T.argmin(1); tb = T[np.arange(n), best]
    Ts = np.sort(T, 1)
    hit = np.isfinite(tb)
    with np.errstate(invalid="ignore"):
        gap = np.where(np.isfinite(Ts[:, 1]), Ts[:, 1] - Ts[:, 0], np.inf)
    clear = hit & (gap > 1e-4 * np.maximum(1.0, Ts[:, 0])) & (edge_margin[np.arange(n), best] > 1e-5)
    # misses must be clear too: no primitive may be missed only by a hair -- inside a closed box nearly every ray hits
    got = w.trace_closest(rays)
    assert clear.sum() > 0.7 * n          # the box is open towards the camera: a fifth of the rays leave through the front
    assert np.array_equal(got["prim"][clear], best[clear].astype(np.uint32))
    np.testing.assert_allclose(got["t"][clear], tb[clear], rtol=2e-5, atol=1e-4)   # f32 coordinates of magnitude 500
    miss = ~hit
    assert np.all(got["prim"][miss & (edge_margin.min(1) > 1e-5)] == 0xFFFFFFFF)


def test_any_hit_consistent_with_closest(make_scene, oracle):
    w = oracle.OracleWorld(make_scene("cornel_box.pbrt"))
    rays = _scene_bbox_rays(20000, seed=11)
    rays[:, 3] = 300.0
    occluded = w.trace_any(rays)
    closest = w.trace_closest(rays)
    assert np.array_equal(occluded.astype(bool), closest["prim"] != 0xFFFFFFFF)


def test_product_bvh_builder_matches_oracle(cpbrt, make_scene, oracle):
    """a8: primitive order, topology, axes and bounds of the exact SAH build -- product host
    builder (cpbrt_host_build_bvh) vs the oracle's restatement of bvh.rs:24-124."""
    for name in ("cornel_box.pbrt", "sphere.pbrt"):
        scene = make_scene(name)
        w = oracle.OracleWorld(scene)
        n = int(scene.desc.n_primitives)
        order = np.zeros(n, np.uint32)
        cap = 2 * n
        arr = dict(is_leaf=np.zeros(cap, np.uint8), axis=np.zeros(cap, np.uint8), a=np.zeros(cap, np.uint32),
                   b=np.zeros(cap, np.uint32), bound=np.zeros((cap, 6), np.float32))
        dump = cpbrt.abi.BvhDump(capacity=cap, **{k: v.ctypes.data_as(C.POINTER({"is_leaf": C.c_uint8, "axis": C.c_uint8, "a": C.c_uint32, "b": C.c_uint32, "bound": C.c_float}[k])) for k, v in arr.items()})
        assert cpbrt.lib.cpbrt_host_build_bvh(scene.desc_ptr, order.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(dump)) == 0
        ref = w.bvh_nodes()
        assert np.array_equal(order, w.bvh_order())
        assert sorted(order) == list(range(n))
        nn = int(dump.n_nodes)
        assert nn == len(ref["a"]) and int(dump.root) == ref["root"] == nn - 1
        for k in ("is_leaf", "axis", "a", "b"):
            assert np.array_equal(arr[k][:nn], ref[k]), k
        assert np.array_equal(arr["bound"][:nn].view(np.uint32), ref["bound"].view(np.uint32))


def _synthetic_mesh_scene(nx, nz, seed=3, res=64):
    """Small height-field trianglemesh in the spirit of config C4."""
    rng = np.random.RandomState(seed)
    xs, zs = np.linspace(0, 556, nx + 1), np.linspace(0, 559, nz + 1)
    h = 40 * rng.rand(nz + 1, nx + 1)
    P = np.stack([np.tile(xs, nz + 1), h.ravel(), np.repeat(zs, nx + 1)], 1)
    idx = []
    for j in range(nz):
        for i in range(nx):
            v = j * (nx + 1) + i
            idx += [v, v + 1, v + nx + 2, v + nx + 2, v + nx + 1, v]
    return ('LookAt 278 273 -800 278 273 0 0 1 0\nCamera "perspective" "float fov" [37.5]\nSampler "halton" "integer pixelsamples" [2]\n'
            'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\nIntegrator "path" "integer maxdepth" [5]\nWorldBegin\n'
            'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [17 12 4]\nShape "trianglemesh" "integer indices" [0 1 2 2 3 0] '
            '"point P" [343 548.7 227 343 548.7 332 213 548.7 332 213 548.7 227]\nAttributeEnd\n'
            'Material "matte" "rgb Kd" [0.5 0.5 0.5]\nShape "trianglemesh" "integer indices" [%s] "point P" [%s]\nWorldEnd\n'
            % (res, res, " ".join(map(str, idx)), " ".join("%.6g" % v for v in P.ravel())))


def test_product_bvh_builder_matches_oracle_on_mesh(cpbrt, oracle):
    """Same as above on a grid mesh: massive key ties (every quad column shares a centre x) and
    2-primitive leaves (quad halves share a bbox, Q7)."""
    scene = cpbrt.Scene.parse(_synthetic_mesh_scene(40, 30))
    n = int(scene.desc.n_primitives)
    assert n == 2 * 40 * 30 + 2
    w = oracle.OracleWorld(scene)
    order = np.zeros(n, np.uint32)
    assert cpbrt.lib.cpbrt_host_build_bvh(scene.desc_ptr, order.ctypes.data_as(C.POINTER(C.c_uint32)), None) == 0
    assert np.array_equal(order, w.bvh_order())
    nodes = w.bvh_nodes()
    leaf_sizes = (nodes["b"] - nodes["a"])[nodes["is_leaf"] == 1]
    assert (leaf_sizes == 2).sum() > 100 and leaf_sizes.max() <= 4


def _triangle_soup_scene(tris):
    """One trianglemesh of independent triangles (n, 3, 3) plus the usual light quad."""
    tris = np.asarray(tris, np.float32)
    n = tris.shape[0]
    return ('LookAt 278 273 -800 278 273 0 0 1 0\nCamera "perspective" "float fov" [37.5]\nSampler "halton" "integer pixelsamples" [2]\n'
            'Film "image" "integer xresolution" [32] "integer yresolution" [32]\nIntegrator "path" "integer maxdepth" [5]\nWorldBegin\n'
            'AttributeBegin\nAreaLightSource "diffuse" "rgb L" [17 12 4]\nShape "trianglemesh" "integer indices" [0 1 2 2 3 0] '
            '"point P" [343 548.7 227 343 548.7 332 213 548.7 332 213 548.7 227]\nAttributeEnd\n'
            'Material "matte" "rgb Kd" [0.5 0.5 0.5]\nShape "trianglemesh" "integer indices" [%s] "point P" [%s]\nWorldEnd\n'
            % (" ".join(map(str, range(3 * n))), " ".join("%.9g" % v for v in tris.ravel())))


def _sah_cases():
    rng = np.random.RandomState(11)
    soup = rng.rand(700, 1, 3) * 500 + (rng.rand(700, 3, 3) - 0.5) * rng.choice([0.5, 8.0, 120.0], (700, 1, 1))
    same = np.tile(np.array([[[1, 1, 1], [2, 1, 1], [1, 2, 1]]], np.float32), (9, 1, 1))             # identical boxes: every cost ties
    flat = np.tile(np.array([[[0, 0, 0], [1, 0, 0], [0, 0, 1]]], np.float32), (12, 1, 1))                # cost 12.125 > 12: one leaf of 12
    line = np.zeros((17, 3, 3), np.float32); line[:, :, 1] = np.arange(17, dtype=np.float32)[:, None]     # zero-area parents: every cost is NaN, the first candidate is taken
    signed = rng.rand(64, 3, 3).astype(np.float32) - 0.5; signed[::4, :, 0] = 0.0; signed[1::4, :, 0] = -0.0   # +0 / -0 centres
    return [("mesh 24x17", None, _synthetic_mesh_scene(24, 17, seed=5)), ("mesh 3x70", None, _synthetic_mesh_scene(3, 70, seed=9)),
            ("soup", soup, None), ("identical", same, None), ("degenerate plane", flat, None), ("points on a line", line, None),
            ("signed zeros", signed, None)]


@pytest.mark.parametrize("case", range(7))
def test_bvh_build_against_independent_numpy_restatement(cpbrt, oracle, case):
    """a8, second pin: oracle/sah_numpy.py restates bvh.rs:24-124 in numpy float32 from the Rust alone
    (own bounds, own sort, the swap loop literally).  Primitive order, node numbering, split axes,
    child links and parent bounds of the C++ oracle and of the product's host builder must equal it
    bit for bit -- on ties, NaN / infinite costs (zero-area parents) and signed-zero keys too."""
    name, tris, text = _sah_cases()[case]
    scene = cpbrt.Scene.parse(text if text is not None else _triangle_soup_scene(tris))
    d = scene.desc
    n = int(d.n_primitives)
    assert all(d.primitives[i].shape.kind == cpbrt.abi.SHAPE_TRIANGLE for i in range(n)), name
    P = np.array([[list(d.triangles[d.primitives[i].shape.data].p[k]) for k in range(3)] for i in range(n)], np.float32)
    sah = _load_oracle_py("sah_numpy")
    tree = sah.build(*sah.triangle_bounds(P))
    w = oracle.OracleWorld(scene)
    ref = w.bvh_nodes()
    assert np.array_equal(tree.order, w.bvh_order()), name
    assert tree.root == ref["root"] and len(tree.a) == len(ref["a"]), name
    for k in ("is_leaf", "a", "b"):
        assert np.array_equal(np.array(getattr(tree, k), np.int64), ref[k].astype(np.int64)), (name, k)
    parents = np.array(tree.is_leaf) == 0
    assert np.array_equal(np.array(tree.axis)[parents], ref["axis"][parents]), name
    assert np.array_equal(np.array(tree.bound, np.float32)[parents].view(np.uint32), ref["bound"][parents].view(np.uint32)), name
    order = np.zeros(n, np.uint32)
    assert cpbrt.lib.cpbrt_host_build_bvh(scene.desc_ptr, order.ctypes.data_as(C.POINTER(C.c_uint32)), None) == 0
    assert np.array_equal(order, tree.order), name


def _load_oracle_py(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(os.path.dirname(__file__), "..", "oracle", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case", [0, 2])
def test_bvh_search_against_independent_numpy_restatement(cpbrt, oracle, case):
    """a9-a12, second pin: oracle/trace_numpy.py restates the far-first search, the loose slab test and the
    watertight triangle test in scalar np.float32 from the Rust alone.  On random rays, camera rays and rays
    aimed exactly at shared vertices and edge midpoints, the C++ oracle must report the same primitive, the
    same bits of t, the same occlusion answer -- and the same number of Parent visits and primitive tests,
    i.e. the same walk."""
    name, tris, text = _sah_cases()[case]
    scene = cpbrt.Scene.parse(text if text is not None else _triangle_soup_scene(tris))
    d = scene.desc
    n = int(d.n_primitives)
    P = np.array([[list(d.triangles[d.primitives[i].shape.data].p[k]) for k in range(3)] for i in range(n)], np.float32)
    sah, trace = _load_oracle_py("sah_numpy"), _load_oracle_py("trace_numpy")
    search = trace.Searcher(sah.build(*sah.triangle_bounds(P)), P)
    lo, hi = P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0)
    rays = [lcg_rays(1200, lo, hi, seed=21 + case)]
    w = oracle.OracleWorld(scene)
    rays.append(w.camera_rays(0)[::2][:400])
    rng = np.random.RandomState(4 + case)
    pick = rng.randint(0, n, 300)
    eye = np.array([278, 900, -300], np.float32)
    aimed = np.zeros((600, 8), np.float32)
    aimed[:, 0:3] = eye; aimed[:, 3] = np.finfo(np.float32).max
    aimed[:300, 4:7] = P[pick, rng.randint(0, 3, 300)] - eye                                # through a vertex
    aimed[300:, 4:7] = (P[pick, 0] + P[pick, 1]) * np.float32(0.5) - eye                        # through an edge midpoint
    rays.append(aimed)
    rays = np.concatenate(rays)
    ref, ref_count = w.trace_closest(rays, count=True, threads=1)
    prim = np.full(len(rays), 0xFFFFFFFF, np.uint32); t = np.zeros(len(rays), np.float32)
    for i, r in enumerate(rays):
        hit, th = search.closest(r[0:3], r[4:7], r[3])
        if hit is not None:
            prim[i], t[i] = hit, th
    assert (prim != 0xFFFFFFFF).sum() > len(rays) // 3, name
    assert np.array_equal(prim, ref["prim"]), name
    assert np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32)), name
    assert (search.visits, search.tests) == (int(ref_count[0]), int(ref_count[1])), name
    # occlusion: segments from the hit points back towards the eye (shortened), plus the raw rays
    seg = rays.copy(); seg[:, 3] = np.where(prim != 0xFFFFFFFF, t * np.float32(1.5), np.float32(50.0))
    both = np.concatenate([rays[:400], seg[::2]])
    search.visits = search.tests = 0
    occluded = np.array([search.any_hit(r[0:3], r[4:7], r[3]) for r in both], np.uint8)
    ref_any, any_count = w.trace_any(both, count=True, threads=1)
    assert np.array_equal(occluded, ref_any), name
    assert (search.visits, search.tests) == (int(any_count[0]), int(any_count[1])), name


def _triangle_cornell_text(variant, **kw):
    """"full": scenes/cornel_box.pbrt as it is (the headline scene: 32 triangles, a glass ball under a Translate, a
    mirror block).  Otherwise the ball is taken out; "matte": the mirror block made matte too; "specular": the mirror
    block kept and the short block made of glass."""
    text = scene_text("cornel_box.pbrt", **kw)
    if variant == "full":
        return text
    ball = 'Translate 188 247.5 176.5\nMaterial "glass"\nShape "sphere" "float radius" [82.5]\n'
    short_block = 'ObjectBegin "ball_and_box"\nMaterial "matte"'
    assert ball in text and 'Material "mirror"' in text and short_block in text
    text = text.replace(ball, "")
    if variant == "matte":
        return text.replace('Material "mirror"', 'Material "matte"')
    return text.replace(short_block, 'ObjectBegin "ball_and_box"\nMaterial "glass"')


def _independent_front(scene):
    """halton_value / camera_ray_at for oracle/path_numpy.render from oracle/front_numpy.py: nothing of the C++ oracle."""
    front = _load_oracle_py("front_numpy")
    d = scene.desc
    perm = np.ctypeslib.as_array(d.sampler.permutation, shape=(cpbrt_perm_len(),))
    cam = front.PerspectiveCamera(list(d.camera.raster_to_camera), list(d.camera.camera_to_world) if d.camera.has_camera_to_world else None,
                                  d.camera.lens_radius if d.camera.has_lens else None, d.camera.focal_distance if d.camera.has_lens else None)

    def halton_value(x, y, s):
        smp = front.HaltonSampler(int(d.sampler.spp), (int(d.film_width), int(d.film_height)), perm)
        smp.set_pixel(x, y)
        for _ in range(s):
            smp.next_sample()

        def value(dim):
            smp.dim = dim
            return smp.get_sample()
        return value
    return halton_value, cam.generate_ray


def cpbrt_perm_len():
    return 3682913      # CPBRT_PERM_TABLE_LEN: the sum of the first 1000 primes


@pytest.mark.parametrize("size", [(900, 900, 1024), (1920, 1080, 1024), (9, 4, 16)])
def test_sampler_and_camera_against_independent_numpy_restatement(cpbrt, oracle, size):
    """a2-a4, second pin: oracle/front_numpy.py (Python integers + np.float32 from halton.rs / perspective.rs alone):
    sampler set-up including the wrapped Bezout coefficient, pixel indices, radical inverses of dimensions 0..999 at
    the BASELINE film sizes, and camera rays under the LookAt transform -- all bit-equal to the C++ oracle."""
    width, height, spp = size
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=width, yres=height, spp=spp), base_dir=SCENES)
    d = scene.desc
    front = _load_oracle_py("front_numpy")
    w = oracle.OracleWorld(scene)
    smp = front.HaltonSampler(spp, (width, height), np.ctypeslib.as_array(d.sampler.permutation, shape=(cpbrt_perm_len(),)))
    assert (smp.scale, smp.exp, smp.mult_inverse) == (list(d.sampler.scale), list(d.sampler.exp), list(d.sampler.mult_inverse))
    rng = np.random.RandomState(1)
    dims = np.arange(0, 1000, 7, dtype=np.uint32)
    for _ in range(25):
        x, y, k = int(rng.randint(width)), int(rng.randint(height)), int(rng.randint(spp))
        smp.set_pixel(x, y)
        for _j in range(k):
            smp.next_sample()
        idx, val = w.halton(np.full(len(dims), x, np.uint32), np.full(len(dims), y, np.uint32), np.full(len(dims), k, np.uint32), dims)
        assert int(idx[0]) == smp.index
        mine = []
        for dim in dims:
            smp.dim = int(dim)
            mine.append(smp.get_sample())
        assert np.array_equal(np.array(mine, np.float32).view(np.uint32), val.view(np.uint32))
    halton_value, camera_ray_at = _independent_front(scene)
    for _ in range(200):
        fx, fy = np.float32(rng.rand() * width), np.float32(rng.rand() * height)
        o, dd = camera_ray_at(fx, fy)
        r = w.camera_ray_at(fx, fy)
        assert np.array_equal(np.array(o + dd, np.float32).view(np.uint32), np.concatenate([r[0:3], r[4:7]]).view(np.uint32))


@pytest.mark.parametrize("variant", ["matte", "specular", "full"])
def test_path_integrator_against_independent_numpy_restatement(cpbrt, oracle, variant):
    """a1, a3, a5, a7, a16, a18, second pin: oracle/path_numpy.py restates the render loop (film-point drift, NaN drop,
    mean), the sampler front, PathIntegrator::li, uniform_sample_one_light with both MIS branches, the area light, the
    triangle sampling, the matte BSDF, the mirror and glass delta lobes (Fresnel, refraction, the luminance-weighted
    choice between them, emission seen through specular bounces), the sphere under a TransformShape (f64 quadratic,
    the hit carried back through ShapePoint::apply with its replaced error bound) and Russian roulette in scalar
    np.float32 from the Rust alone, over the numpy BVH.
    Every pixel of the C++ oracle's image must equal it bit for bit, at every path depth, with the same ray counts.
    (Its first run differed in 3 of 576 pixels by one ulp: the new reading had missed that Normal3f::from normalises
    an already normalised vector -- the C++ oracle had it.)"""
    scene = cpbrt.Scene.parse(_triangle_cornell_text(variant, xres=22, yres=22, spp=3), base_dir=SCENES)
    kinds = {int(scene.desc.materials[scene.desc.primitives[i].material].type) for i in range(int(scene.desc.n_primitives)) if scene.desc.primitives[i].material >= 0}
    assert kinds == ({0} if variant == "matte" else {0, 1, 2})
    assert int(scene.desc.n_spheres) == (1 if variant == "full" else 0)
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    assert np.array_equal(mine.search.tree.order, w.bvh_order())

    def halton_value(x, y, s):
        dims = np.arange(96, dtype=np.uint32)
        _, v = w.halton(np.full(96, x, np.uint32), np.full(96, y, np.uint32), np.full(96, s, np.uint32), dims)
        return lambda dim: v[dim]

    def camera_ray_at(fx, fy):
        r = w.camera_ray_at(fx, fy)
        return r[0:3], r[4:7]

    if variant == "full":      # scene description in, image out: nothing of the C++ oracle on this side
        halton_value, camera_ray_at = _independent_front(scene)
    for depth, rect in ((5, None), (2, (3, 9, 19, 13)), (0, (0, 0, 22, 4))):
        mine.max_depth = depth
        mine.rays_closest = mine.rays_any = 0
        film = path.render(mine, 22, 22, 3, halton_value, camera_ray_at, rect=rect)
        ref, _, st = w.render(max_depth=depth, rect=rect or (0, 0, 0, 0))
        assert (film > 0).any() and np.isfinite(film).all()
        assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"]), depth
        assert np.array_equal(film.view(np.uint32), ref.view(np.uint32)), depth


def test_directlighting_against_independent_numpy_restatement(cpbrt, oracle):
    """a6, second pin: DirectLightIntegrator::li_depth of oracle/path_numpy.py (the one-light estimate run on all-delta
    surfaces too, the recursion over every delta lobe of the mirror block and the glass ball) on the unmodified
    Cornell box: bit-equal image, equal ray counts."""
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=22, yres=22, spp=3, integrator=("directlighting", 5)), base_dir=SCENES)
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)

    def halton_value(x, y, s):
        dims = np.arange(256, dtype=np.uint32)
        _, v = w.halton(np.full(256, x, np.uint32), np.full(256, y, np.uint32), np.full(256, s, np.uint32), dims)
        return lambda dim: v[dim]

    def camera_ray_at(fx, fy):
        r = w.camera_ray_at(fx, fy)
        return r[0:3], r[4:7]

    film = path.render(mine, 22, 22, 3, halton_value, camera_ray_at, integrator="directlighting")
    ref, _, st = w.render()
    assert (film > 0).any() and np.isfinite(film).all()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


def test_sphere_scene_against_independent_numpy_restatement(cpbrt, oracle):
    """BASELINE config C2 (scenes/sphere.pbrt: matte spheres under Translate, a point light, a distant light, a
    spherical emitter sampled by its cone, directlighting, a translated camera) through oracle/{front,sah,trace,path}_numpy.py
    alone: bit-equal to the C++ oracle, equal ray counts."""
    scene = cpbrt.Scene.parse(scene_text("sphere.pbrt", xres=48, yres=27, spp=3), base_dir=SCENES)
    assert [int(scene.desc.lights[k].type) for k in range(int(scene.desc.n_lights))] == [1, 0, 2]
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    assert np.array_equal(mine.search.tree.order, w.bvh_order())
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, 48, 27, 3, halton_value, camera_ray_at, integrator="directlighting")
    ref, _, st = w.render()
    assert (film > 0).any() and np.isfinite(film).all()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("config", ["C3", "C4"])
def test_textured_and_mesh_scenes_against_independent_numpy_restatement(cpbrt, oracle, config):
    """BASELINE configs C3 (the Cornell box with box-texture.png as Kd: nearest-texel look-up through the per-triangle
    uv) and C4 in small (a height-field trianglemesh under the light quad: two-primitive leaves, massive key ties in
    the build) through oracle/{front,sah,trace,path}_numpy.py alone: bit-equal to the C++ oracle, equal ray counts."""
    from conftest import c3_text
    if config == "C3":
        text = c3_text()
        for key, old, new in (("xresolution", 1920, 40), ("yresolution", 1080, 24), ("pixelsamples", 1024, 2)):
            assert '"integer %s" [%d]' % (key, old) in text
            text = text.replace('"integer %s" [%d]' % (key, old), '"integer %s" [%d]' % (key, new))
        width, height, spp = 40, 24, 2
    else:
        text, (width, height, spp) = _synthetic_mesh_scene(14, 11, seed=8, res=20), (20, 20, 2)
    scene = cpbrt.Scene.parse(text, base_dir=SCENES)
    assert (int(scene.desc.film_width), int(scene.desc.film_height), int(scene.desc.sampler.spp)) == (width, height, spp)
    if config == "C3":
        assert any(int(scene.desc.textures[k].kind) == 1 for k in range(int(scene.desc.n_textures)))
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    assert np.array_equal(mine.search.tree.order, w.bvh_order())
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, width, height, spp, halton_value, camera_ray_at)
    ref, _, st = w.render()
    assert (film > 0).any() and np.isfinite(film).all()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("name", ["mix", "mix_black_amount", "plastic", "plastic_rough", "translucent", "uber"])
def test_general_materials_against_independent_numpy_restatement(cpbrt, oracle, name):
    """f1, second pin: plastic / uber / translucent / mix, Trowbridge-Reitz sampling and evaluation, microfacet reflection
    and transmission (with the trait-default pdf the reference leaves in place), Lambertian transmission, Oren-Nayar,
    BSDF::mix with its scaled lobes -- oracle/path_numpy.py against the C++ oracle on the Cornell box with its white
    surfaces replaced: bit-equal image, equal ray counts."""
    from test_gpu_parity import GENERAL_MATERIALS, _cornell_with_white_material
    scene = cpbrt.Scene.parse(_cornell_with_white_material(GENERAL_MATERIALS[name], xres=18, yres=18, spp=2), base_dir=SCENES)
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, 18, 18, 2, halton_value, camera_ray_at)
    ref, _, st = w.render()
    assert (film > 0).any() and st["nan_dropped"] == 0
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


def test_thin_lens_against_independent_numpy_restatement(cpbrt, oracle):
    """f4, second pin: LensCamera::generate_ray (two numbers drawn after the film offset, lens point on the concentric
    disk, refocused through the focal plane, then the camera transform) on the Cornell box: bit-equal image."""
    text = scene_text("cornel_box.pbrt", xres=20, yres=20, spp=3)
    assert '"float fov"' in text
    scene = cpbrt.Scene.parse(text.replace('"float fov"', '"float lensradius" [12] "float focaldistance" [900] "float fov"', 1), base_dir=SCENES)
    assert scene.desc.camera.has_lens
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, 20, 20, 3, halton_value, camera_ray_at, camera_draws=True)
    ref, _, st = w.render()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


INFINITE_CASES = ["cornell constant path", "cornell envmap path", "cornell envmap directlighting", "furnace matte path",
                  "furnace matte directlighting", "furnace glass"]


@pytest.mark.parametrize("case", INFINITE_CASES)
def test_infinite_area_light_against_independent_numpy_restatement(cpbrt, oracle, case):
    """f1, second pin: InfiniteAreaLight of oracle/path_numpy.py -- the sin((row + 0.5) / rows) table weights, Distribution2D
    with its first coordinate over the rows, the lat-long mapping (atan2f / acosf), le_out_scene for escaped camera and
    specular rays, the escaped-ray MIS branch -- in the Cornell box next to the area light (constant and box-texture.png
    environments) and in the open furnace scenes where every camera ray can escape: bit-equal image, equal ray counts."""
    import re
    width = height = 16
    spp = 2
    if case.startswith("cornell"):
        light = 'LightSource "infinite"' + (' "string mapname" "box-texture.png"' if "envmap" in case else "")
        text = scene_text("cornel_box.pbrt", xres=width, yres=height, spp=spp).replace("WorldBegin\n", "WorldBegin\n" + light + "\n", 1)
        if "directlighting" in case:
            text = re.sub(r'Integrator "path"[^\n]*\n(\s*"integer maxdepth"[^\n]*\n)?', 'Integrator "directlighting" "integer maxdepth" [1]\n', text)
    elif case == "furnace glass":
        text = specular_furnace_scene_text("glass", spp=spp, res=width)
    else:
        text = furnace_scene_text(case.split()[-1], 0.75, spp=spp, res=width)
    scene = cpbrt.Scene.parse(text, base_dir=SCENES)
    assert any(int(scene.desc.lights[k].type) == 3 for k in range(int(scene.desc.n_lights)))
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, width, height, spp, halton_value, camera_ray_at, integrator="directlighting" if "directlighting" in case else "path")
    ref, _, st = w.render()
    assert (film > 0).any() and np.isfinite(film).all()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


def test_instances_under_transforms_against_independent_numpy_restatement(cpbrt, oracle):
    """ObjectInstance under Translate / Rotate / Scale (src/scene.rs:119-140, src/primitive.rs:115-130): triangles and
    spheres behind one and two nested TransformShapes -- rays through both inverses, hits back through both
    ShapePoint::apply steps, bounds as unions of transformed corners -- plus an image texture on a mesh with its own uv
    and on a sphere (Sphere::calc_uv: atan2f / acosf): bit-equal image, equal BVH order and ray counts."""
    width = height = 22
    spp = 2
    text = ('LookAt 0 1.5 -6 0 0.5 0 0 1 0\nCamera "perspective" "float fov" [45]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\nIntegrator "path" "integer maxdepth" [4]\nWorldBegin\n'
            'Texture "boxtex" "spectrum" "imagemap" "string filename" "box-texture.png"\n'
            'LightSource "infinite"\n'
            'AttributeBegin\nTranslate 0 4 0\nLightSource "point" "rgb I" [30 30 30]\nAttributeEnd\n'
            'Material "matte" "rgb Kd" [0.6 0.6 0.6]\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-5 0 -5 5 0 -5 5 0 5 -5 0 5]\n'
            'ObjectBegin "thing"\n'
            'Material "matte" "texture Kd" "boxtex"\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 0 0 1 0 0 1 1.5 0.3 -1 1.5 0.3] "float uv" [0 0 1 0 1 1 0 1]\n'
            'Translate 0 0.8 -1\n'
            'Shape "sphere" "float radius" [0.5]\n'
            'Material "glass"\nTranslate 1.2 0 0\nShape "sphere" "float radius" [0.4]\n'
            'ObjectEnd\n'
            'AttributeBegin\nTranslate -1.5 0 1\nRotate 35 0 1 0\nScale 1.2 0.8 1\nObjectInstance "thing"\nAttributeEnd\n'
            'AttributeBegin\nTranslate 1.8 0.2 0\nRotate -20 0 0 1\nObjectInstance "thing"\nAttributeEnd\n'
            'WorldEnd\n' % (spp, width, height))
    scene = cpbrt.Scene.parse(text, base_dir=SCENES)
    d = scene.desc
    depth = []
    for i in range(int(d.n_primitives)):
        x, k = int(d.primitives[i].shape.xform), 0
        while x >= 0:
            x, k = int(d.transforms[x].outer), k + 1
        depth.append((int(d.primitives[i].shape.kind), k))
    assert (0, 1) in depth and (1, 2) in depth          # a triangle behind one wrapper, a sphere behind two
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(d, cpbrt.abi, sah, trace)
    assert np.array_equal(mine.search.tree.order, w.bvh_order())
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, width, height, spp, halton_value, camera_ray_at)
    ref, _, st = w.render()
    assert (film > 0).any() and np.isfinite(film).all()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("integrator", ["path", "directlighting"])
@pytest.mark.parametrize("transform", ["", "Translate 0.5 -0.3 1\nScale 1 1.3 0.9\n"])
def test_enclosing_sphere_emitter_against_independent_numpy_restatement(cpbrt, oracle, integrator, transform):
    """Shading points INSIDE a spherical emitter (Shape::default_sample_by_point over Sphere::sample, the area-measure
    pdf with its sign, default_by_point_pdf on the MIS side), the emitter bare (no TransformShape at all) and behind a
    Translate + Scale: bit-equal image, equal ray counts."""
    width = height = 18
    spp = 2
    text = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\nIntegrator "%s" "integer maxdepth" [3]\nWorldBegin\n'
            'AttributeBegin\n%sAreaLightSource "diffuse" "rgb L" [1 0.9 0.8]\nShape "sphere" "float radius" [30]\nAttributeEnd\n'
            'Material "matte" "rgb Kd" [0.5 0.6 0.7]\nAttributeBegin\nTranslate 0 0 5\nShape "sphere" "float radius" [1]\nAttributeEnd\n'
            'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-3 -1 2 3 -1 2 3 -1 9 -3 -1 9]\nWorldEnd\n'
            % (spp, width, height, integrator, transform))
    scene = cpbrt.Scene.parse(text)
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, width, height, spp, halton_value, camera_ray_at, integrator=integrator)
    ref, _, st = w.render()
    assert (film > 0).any() and np.isfinite(film).all()
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))


def test_random_scenes_against_independent_numpy_restatement(cpbrt, oracle):
    """A slice of tools/independent_restatement_fuzz.py: ten random scenes (triangles and spheres bare, transformed and
    instanced, every material and light type, both integrators, thin lens or not) -- bit-equal images, equal BVH order
    and ray counts.  500 scenes of the same generator: profiles/r02_independent_restatement_fuzz.txt."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools", "independent_restatement_fuzz.py")
    out = subprocess.run([sys.executable, tool, "--scenes", "10", "--seed", "7"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    summary = json.loads(out.stdout.strip().splitlines()[-1])
    assert summary["scenes"] >= 8 and summary["mismatching_scenes"] == [] and summary["pixels_bit_equal"] == summary["pixels"]


def test_degenerate_scenes_against_independent_numpy_restatement(cpbrt, oracle):
    """An empty world, only an environment, a one-pixel film, no lights at all, everything missed: bit-equal images."""
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [%d]\n'
            'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\nIntegrator "path" "integer maxdepth" [3]\nWorldBegin\n')
    cases = [
        (2, 8, 8, 'WorldEnd\n'),
        (2, 8, 8, 'LightSource "infinite"\nWorldEnd\n'),
        (1, 1, 1, 'LightSource "point" "point from" [0 2 0]\nMaterial "matte"\nTranslate 0 0 5\nShape "sphere"\nWorldEnd\n'),
        (3, 16, 4, 'Material "matte"\nTranslate 0 0 5\nShape "sphere" "float radius" [0.5]\nWorldEnd\n'),
        (2, 12, 12, 'LightSource "distant"\nMaterial "mirror"\nTranslate 100 0 5\nShape "sphere"\nWorldEnd\n'),
    ]
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    for spp, width, height, body in cases:
        scene = cpbrt.Scene.parse(head % (spp, width, height) + body, base_dir=SCENES)
        w = oracle.OracleWorld(scene)
        mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
        halton_value, camera_ray_at = _independent_front(scene)
        film = path.render(mine, width, height, spp, halton_value, camera_ray_at)
        ref, _, st = w.render()
        assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"]), body
        assert np.array_equal(film.view(np.uint32), ref.view(np.uint32)), body


def test_nan_samples_against_independent_numpy_restatement(cpbrt, oracle):
    """The NaN drop of the render loop (render.rs:33-40) and the per-pixel mean over the samples that were kept -- 0 / 0
    where none was.  A glass ball with Kr = Kt = 0 makes them: both delta lobes are black, the luminance table is all
    zero, its pmf NaN, and `beta` is NaN from there on; `l += beta * intersect.le()` then poisons `l` at the next specular
    vertex even off the emitters (le() = None multiplies as black and is still added) -- the line the second reading got
    wrong at first, by adding only on emitters.  Same dropped-sample count, same pixels NaN, every other pixel bit-equal."""
    width = height = 20
    spp = 3
    text = scene_text("cornel_box.pbrt", xres=width, yres=height, spp=spp).replace('Material "glass"', 'Material "glass" "rgb Kr" [0 0 0] "rgb Kt" [0 0 0]')
    scene = cpbrt.Scene.parse(text, base_dir=SCENES)
    sah, trace, path = (_load_oracle_py(n) for n in ("sah_numpy", "trace_numpy", "path_numpy"))
    w = oracle.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    halton_value, camera_ray_at = _independent_front(scene)
    film = path.render(mine, width, height, spp, halton_value, camera_ray_at)
    ref, _, st = w.render()
    assert st["nan_dropped"] > 50 and mine.nan_dropped == st["nan_dropped"]
    assert (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"])
    assert np.isnan(ref).any() and np.array_equal(np.isnan(film), np.isnan(ref))
    assert np.array_equal(film.view(np.uint32)[~np.isnan(ref)], ref.view(np.uint32)[~np.isnan(ref)])


def _adversarial_triangles(kind):
    rng = np.random.RandomState(23)
    if kind in ("scaled up", "scaled down"):
        soup = rng.rand(300, 1, 3) * 500 + (rng.rand(300, 3, 3) - 0.5) * rng.choice([0.5, 8.0, 120.0], (300, 1, 1))
        return (soup * (1e3 if kind == "scaled up" else 1e-3)).astype(np.float32)
    if kind == "coincident":
        # the same quad stacked 0, 1, 2, 4 ... 64 ulp apart along z, twice over with opposite windings
        z = np.float32(300.0)
        layers = [z]
        for step in (0, 1, 1, 2, 4, 8, 16, 32):
            nxt = layers[-1]
            for _ in range(step):
                nxt = np.nextafter(nxt, np.float32(np.inf))
            layers.append(nxt)
        tris = []
        for k, zz in enumerate(layers):
            quad = np.array([[100, 100, zz], [400, 100, zz], [400, 400, zz], [100, 400, zz]], np.float32)
            order = (0, 1, 2, 2, 3, 0) if k % 2 == 0 else (0, 2, 1, 0, 3, 2)
            tris += [quad[list(order[:3])], quad[list(order[3:])]]
        return np.array(tris, np.float32)
    assert kind == "slivers"
    base = rng.rand(200, 1, 3) * 400 + 50
    long_edge = (rng.rand(200, 1, 3) - 0.5) * 300
    thin = (rng.rand(200, 1, 3) - 0.5) * rng.choice([1e-4, 1e-2, 1.0], (200, 1, 1))
    return np.concatenate([base, base + long_edge, base + long_edge * 0.5 + thin], axis=1).astype(np.float32)


@pytest.mark.parametrize("kind", ["scaled up", "scaled down", "coincident", "slivers"])
def test_adversarial_geometry_against_independent_numpy_restatement(cpbrt, oracle, kind):
    """Coordinates of 1e5 and 1e-1, the same quad stacked 0 to 64 ulp apart with alternating windings, slivers down to
    1e-4 wide: BVH order, closest primitive, the bits of t, occlusion and the visit / test counts of the C++ oracle
    against the numpy restatement -- where ties in t are decided by traversal order alone."""
    tris = _adversarial_triangles(kind)
    look = {"scaled up": "LookAt 278000 273000 -800000 278000 273000 0 0 1 0", "scaled down": "LookAt 0.278 0.273 -0.8 0.278 0.273 0 0 1 0"}
    scene = cpbrt.Scene.parse(_triangle_soup_scene(tris).replace("LookAt 278 273 -800 278 273 0 0 1 0", look.get(kind, "LookAt 278 273 -800 278 273 0 0 1 0")))
    d = scene.desc
    n = int(d.n_primitives)
    assert n >= len(tris)                                            # the clipper kept the soup
    P = np.array([[list(d.triangles[d.primitives[i].shape.data].p[k]) for k in range(3)] for i in range(n)], np.float32)
    sah, trace = _load_oracle_py("sah_numpy"), _load_oracle_py("trace_numpy")
    search = trace.Searcher(sah.build(*sah.triangle_bounds(P)), P)
    w = oracle.OracleWorld(scene)
    assert np.array_equal(search.tree.order, w.bvh_order()), kind
    own = P[n - len(tris):]                                         # without the light quad of the scene template
    lo, hi = own.reshape(-1, 3).min(0), own.reshape(-1, 3).max(0)
    pad = (hi - lo) * 0.2 + 1e-3
    rays = [lcg_rays(500, lo - pad, hi + pad, seed=41)]
    rng = np.random.RandomState(5)
    pick = rng.randint(0, len(own), 300)
    eye = (lo + hi) * 0.5 + np.array([0.3, 0.2, -3.0], np.float32) * (hi - lo + 1.0)
    aimed = np.zeros((600, 8), np.float32)
    aimed[:, 0:3] = eye; aimed[:, 3] = np.finfo(np.float32).max
    aimed[:300, 4:7] = own[pick, rng.randint(0, 3, 300)] - eye                       # through a vertex
    aimed[300:, 4:7] = own[pick].mean(axis=1) - eye                                    # through a centroid
    rays = np.concatenate(rays + [aimed]).astype(np.float32)
    ref, ref_count = w.trace_closest(rays, count=True, threads=1)
    prim = np.full(len(rays), 0xFFFFFFFF, np.uint32); t = np.zeros(len(rays), np.float32)
    for i, r in enumerate(rays):
        hit, th = search.closest(r[0:3], r[4:7], r[3])
        if hit is not None:
            prim[i], t[i] = hit, th
    assert (prim != 0xFFFFFFFF).sum() > 100, kind
    assert np.array_equal(prim, ref["prim"]), kind
    assert np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32)), kind
    assert (search.visits, search.tests) == (int(ref_count[0]), int(ref_count[1])), kind
    seg = rays.copy(); seg[:, 3] = np.where(prim != 0xFFFFFFFF, t * np.float32(1.0000001), np.float32(50.0))
    search.visits = search.tests = 0
    occluded = np.array([search.any_hit(r[0:3], r[4:7], r[3]) for r in seg], np.uint8)
    ref_any, any_count = w.trace_any(seg, count=True, threads=1)
    assert np.array_equal(occluded, ref_any), kind
    assert (search.visits, search.tests) == (int(any_count[0]), int(any_count[1])), kind


def test_render_is_deterministic_and_thread_count_independent(make_scene, oracle):
    scene = make_scene("cornel_box.pbrt", xres=40, yres=40, spp=2)
    w = oracle.OracleWorld(scene)
    a, _, sa = w.render(threads=1)
    b, _, sb = w.render(threads=8)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and sa == sb
    assert np.isfinite(a).all() and a.mean() > 0.01


def test_golden_vectors(make_scene, oracle):
    """Committed oracle outputs (tests/golden/make_golden.py): guards the oracle itself against
    accidental change, here and on the GPU box."""
    g = np.load(os.path.join(GOLDEN, "cornell_golden.npz"))
    scene = make_scene("cornel_box.pbrt", xres=int(g["xres"]), yres=int(g["yres"]), spp=int(g["spp"]))
    w = oracle.OracleWorld(scene)
    assert np.array_equal(w.bvh_order(), g["bvh_order"])
    rays = w.camera_rays(0)
    assert np.array_equal(rays.view(np.uint32), g["camera_rays"].view(np.uint32))
    hits = w.trace_closest(rays)
    assert np.array_equal(hits["prim"], g["first_hit_prim"])
    assert np.array_equal(hits["t"].view(np.uint32), g["first_hit_t"].view(np.uint32))
    idx, val = w.halton(g["h_px"], g["h_py"], g["h_sample"], g["h_dim"])
    assert np.array_equal(idx, g["h_index"])
    assert np.array_equal(val.view(np.uint32), g["h_value"].view(np.uint32))
    film, _, st = w.render(threads=8)
    # libm (sin/cos/atan2/acos/pow) may differ by an ulp between machines: compare loosely
    np.testing.assert_allclose(film, g["film"], rtol=2e-3, atol=2e-4)
    assert abs(st["rays_closest"] - int(g["rays_closest"])) <= 0.001 * int(g["rays_closest"])
