"""Host parser / scene builder against facts derived from the reference's sources
 -- not produced by a reference run."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENES, scene_text, write_exr


def test_cornell_tables(cpbrt):
    s = cpbrt.Scene.load(os.path.join(SCENES, "cornel_box.pbrt"))
    d = s.desc
    abi = cpbrt.abi
    assert (d.film_width, d.film_height) == (900, 900)
    assert d.integrator == abi.INTEGRATOR_PATH and d.max_depth == 5
    assert d.n_primitives == 33 and d.n_triangles == 32 and d.n_spheres == 1 and d.n_lights == 2
    assert s.film_filename == "image/cornel_test.png"
    prims = [d.primitives[i] for i in range(33)]
    assert [p.light for p in prims[:2]] == [0, 1] and all(p.material < 0 for p in prims[:2])
    assert sum(p.shape.kind == abi.SHAPE_SPHERE for p in prims) == 1
    sphere = [p for p in prims if p.shape.kind == abi.SHAPE_SPHERE][0]
    assert d.materials[sphere.material].type == abi.MAT_GLASS
    t = d.transforms[sphere.shape.xform]
    assert (t.m[3], t.m[7], t.m[11]) == (188.0, 247.5, 176.5) and t.outer == -1
    assert abs(d.sphere_radius[sphere.shape.data] - 82.5) == 0
    assert [d.materials[p.material].type for p in prims[-10:]] == [abi.MAT_MIRROR] * 10
    # spectrum -> RGB through from_sampled (rgb_spectrum.rs:85-96) and the CIE tables
    np.testing.assert_allclose(list(d.lights[0].L), [18.487453, 11.010248, 2.757321], rtol=1e-5)
    kd = lambda m: list(d.textures[d.materials[m].tex[0]].value)
    np.testing.assert_allclose(kd(prims[2].material), [0.582553, 0.043657, 0.044207], rtol=2e-5)
    np.testing.assert_allclose(kd(prims[4].material), [0.100772, 0.376082, 0.078298], rtol=2e-5)
    np.testing.assert_allclose(kd(prims[6].material), [0.886975, 0.696752, 0.667831], rtol=2e-5)
    # HaltonSampler::new(1, (900, 900))
    assert list(d.sampler.scale) == [1024, 2187] and list(d.sampler.exp) == [10, 7]
    assert (d.sampler.mult_inverse[0] * 2187) % 1024 == 1 and (d.sampler.mult_inverse[1] * 1024) % 2187 == 1
    # LookAt 278 273 -800 -> camera-to-world is a pure translation here
    c2w = np.array(list(d.camera.camera_to_world)).reshape(4, 4)
    np.testing.assert_array_equal(c2w, [[1, 0, 0, 278], [0, 1, 0, 273], [0, 0, 1, -800], [0, 0, 0, 1]])


def test_sphere_scene_tables(cpbrt):
    s = cpbrt.Scene.load(os.path.join(SCENES, "sphere.pbrt"))
    d = s.desc
    abi = cpbrt.abi
    assert (d.film_width, d.film_height, d.sampler.spp) == (1600, 900, 8)
    assert d.integrator == abi.INTEGRATOR_DIRECT and d.max_depth == 0
    assert d.n_primitives == 4 and d.n_spheres == 4
    assert [d.lights[i].type for i in range(3)] == [abi.LIGHT_POINT, abi.LIGHT_AREA, abi.LIGHT_DISTANT]
    assert list(d.lights[0].v) == [-1.0, 1.0, 3.0]
    assert list(d.lights[2].v) == [0.0, 0.0, 1.0]           # normalize(to - from), translation does not move a vector
    assert d.primitives[3].light == 1 and d.primitives[3].material == -1
    # only the first top-level transform directive becomes the camera transform (render.rs:65)
    c2w = np.array(list(d.camera.camera_to_world)).reshape(4, 4)
    np.testing.assert_array_equal(c2w[:3, 3], [1, 1, -4])


def test_perspective_works_kat(cpbrt, oracle):
    """src/camera/perspective.rs:88-105: PerspectiveCamera::new(40, (1024, 768))."""
    s = cpbrt.Scene.parse('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
                          'Film "image" "integer xresolution" [1024] "integer yresolution" [768]\nIntegrator "path"\nWorldBegin\nWorldEnd\n')
    d = s.desc
    m = np.array(list(d.camera.camera_to_raster), np.float32).reshape(4, 4)
    p = np.array([0, 0, 1000, 1], np.float32)
    r = np.zeros(4, np.float32)
    for i in range(4):
        r[i] = np.float32(np.float32(np.float32(m[i, 0] * p[0] + m[i, 1] * p[1]) + m[i, 2] * p[2]) + m[i, 3] * p[3])
    assert tuple(r[:3] / r[3]) == (512.0, 384.0, 1.0)
    ray = oracle.OracleWorld(s).camera_ray_at(512.0, 384.0)
    assert tuple(ray[:3]) == (0.0, 0.0, 0.0) and tuple(ray[4:7]) == (0.0, 0.0, 1.0) and ray[3] == np.finfo(np.float32).max
    # the stored pair is a true inverse pair
    inv = np.array(list(d.camera.raster_to_camera), np.float64).reshape(4, 4)
    np.testing.assert_allclose(inv @ m.astype(np.float64), np.eye(4), atol=2e-3)


def test_transform_semantics(cpbrt):
    """CTM composition is new * old and Rotate is the transpose of pbrt's (Q17); instancing
    wraps instead of composing (primitive.rs:115-130)."""
    head = ('Camera "perspective" "float fov" [90]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [64] "integer yresolution" [64]\nIntegrator "path"\nWorldBegin\nMaterial "matte"\n')
    s = cpbrt.Scene.parse(head + 'AttributeBegin\nTranslate 1 2 3\nScale 2 2 2\nShape "sphere"\nAttributeEnd\nWorldEnd\n')
    d = s.desc
    assert d.n_primitives == 1
    t = d.transforms[d.primitives[0].shape.xform]
    m = np.array(list(t.m)).reshape(4, 4)
    # new * old = Scale * Translate: the translation gets scaled
    np.testing.assert_array_equal(m, [[2, 0, 0, 2], [0, 2, 0, 4], [0, 0, 2, 6], [0, 0, 0, 1]])
    s = cpbrt.Scene.parse(head + 'AttributeBegin\nRotate 90 0 0 1\nShape "sphere"\nAttributeEnd\nWorldEnd\n')
    m = np.array(list(s.desc.transforms[0].m)).reshape(4, 4)
    np.testing.assert_allclose(m[:2, :2], [[0, 1], [-1, 0]], atol=1e-6)       # pbrt would give [[0,-1],[1,0]]
    s = cpbrt.Scene.parse(head + 'ObjectBegin "o"\nTranslate 0 0 5\nShape "sphere"\nObjectEnd\n'
                          'AttributeBegin\nTranslate 1 0 0\nObjectInstance "o"\nAttributeEnd\nWorldEnd\n')
    d = s.desc
    assert d.n_primitives == 1
    inner = d.transforms[d.primitives[0].shape.xform]
    assert inner.m[11] == 5.0 and inner.outer >= 0 and d.transforms[inner.outer].m[3] == 1.0


def test_clipper_drops_offscreen_material_primitives(cpbrt):
    """perspective.rs:54-78 tests only the bbox min and max corners; emissive primitives are kept."""
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [64] "integer yresolution" [64]\nIntegrator "path"\nWorldBegin\nMaterial "matte"\n')
    tri = lambda x: ('Shape "trianglemesh" "integer indices" [0 1 2] "point P" [%g -0.1 5 %g 0.1 5 %g 0.1 5]\n' % (x, x + 0.1, x))
    s = cpbrt.Scene.parse(head + tri(0) + tri(100) + 'AttributeBegin\nAreaLightSource "diffuse"\n' + tri(100) + 'AttributeEnd\nWorldEnd\n')
    d = s.desc
    assert d.n_primitives == 2                      # on-screen matte + off-screen emitter; off-screen matte dropped
    assert d.primitives[0].material >= 0 and d.primitives[1].light == 0


def test_image_texture_load(cpbrt):
    """`Texture "name" "spectrum" "imagemap"`: PNG bytes / 255 then the 1.05 inverse gamma (Q19)."""
    from PIL import Image
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [8] "integer yresolution" [8]\nIntegrator "path"\nWorldBegin\n')
    s = cpbrt.Scene.parse(head + 'Texture "boxtex" "spectrum" "imagemap" "string filename" "box-texture.png"\n'
                          'Material "matte" "texture Kd" "boxtex"\nShape "sphere"\nWorldEnd\n', base_dir=SCENES)
    d = s.desc
    tex = d.textures[d.materials[0].tex[0]]
    assert (tex.kind, tex.channels, tex.width, tex.height) == (cpbrt.abi.TEX_IMAGE, 3, 474, 474)
    img = np.asarray(Image.open(os.path.join(SCENES, "box-texture.png")).convert("RGB")).astype(np.float32) / np.float32(255)
    expect = np.where(img <= 0.04045, img / 12.92, ((img + 0.055) / 1.05) ** 2.4)
    got = np.ctypeslib.as_array(d.texels, shape=(int(d.n_texels),))[tex.offset:tex.offset + 474 * 474 * 3].reshape(474, 474, 3)
    np.testing.assert_allclose(got, expect, rtol=2e-5, atol=1e-7)


def test_plymesh_and_include(cpbrt, tmp_path):
    """Binary little-endian PLY meshes (geometry/shape/plymesh.rs:14-131: float x y z nx ny nz u v
    per vertex, `uint8 int` faces) and the `Include` directive (lex.rs:98-105): the same mesh as
    `trianglemesh` text, as a .ply file, and through an included file gives identical triangles."""
    import struct
    rng = np.random.default_rng(3)
    verts = (rng.random((40, 3)).astype(np.float32) * 2 - 1) * np.float32(0.5) + np.array([0, 0, 5], np.float32)
    faces = rng.integers(0, 40, size=(60, 3)).astype(np.int32)
    with open(tmp_path / "m.ply", "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 40\n")
        for name in ("x", "y", "z", "nx", "ny", "nz", "u", "v"):
            f.write(("property float %s\n" % name).encode())
        f.write(b"element face 60\nproperty list uint8 int vertex_indices\nend_header\n")
        for v in verts:
            f.write(struct.pack("<8f", v[0], v[1], v[2], 0.0, 0.0, 1.0, 0.25, 0.75))
        for t in faces:
            f.write(struct.pack("<B3i", 3, *t))
    head = ('Camera "perspective" "float fov" [60]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [32] "integer yresolution" [32]\nIntegrator "path"\nWorldBegin\nMaterial "matte"\n')
    mesh_text = 'Shape "trianglemesh" "integer indices" [%s] "point P" [%s]\n' % (
        " ".join(str(i) for i in faces.ravel()), " ".join(repr(float(x)) for x in verts.ravel()))
    (tmp_path / "inc.pbrt").write_text('Shape "plymesh" "string filename" "m.ply"\n')
    # the Scene objects own the arrays their descriptions point into: keep them alive while comparing
    scene_a = cpbrt.Scene.parse(head + mesh_text + "WorldEnd\n", base_dir=str(tmp_path))
    scene_b = cpbrt.Scene.parse(head + 'Shape "plymesh" "string filename" "m.ply"\nWorldEnd\n', base_dir=str(tmp_path))
    scene_c = cpbrt.Scene.parse(head + 'Include "inc.pbrt"\nWorldEnd\n', base_dir=str(tmp_path))
    a, b, c = scene_a.desc, scene_b.desc, scene_c.desc
    assert a.n_triangles > 0
    ta = np.ctypeslib.as_array(a.triangles, shape=(int(a.n_triangles),)).view(np.uint8).copy()
    for other in (b, c):
        assert other.n_triangles == a.n_triangles and other.n_primitives == a.n_primitives
        assert np.array_equal(np.ctypeslib.as_array(other.triangles, shape=(int(other.n_triangles),)).view(np.uint8), ta)
    # a header cut off inside its property list, and element counts far beyond the file: error codes, promptly
    data = (tmp_path / "m.ply").read_bytes()
    (tmp_path / "cut.ply").write_bytes(data[:data.index(b"property float nz")])
    (tmp_path / "many.ply").write_bytes(data.replace(b"element vertex 40", b"element vertex 4000000000"))
    for name in ("cut.ply", "many.ply"):
        with pytest.raises(cpbrt.CpbrtError, match="truncated"):
            cpbrt.Scene.parse(head + 'Shape "plymesh" "string filename" "%s"\nWorldEnd\n' % name, base_dir=str(tmp_path))
    # a truncated file is an error code, not a crash
    (tmp_path / "bad.ply").write_bytes(data[:len(data) // 2])
    with pytest.raises(cpbrt.CpbrtError):
        cpbrt.Scene.parse(head + 'Shape "plymesh" "string filename" "bad.ply"\nWorldEnd\n', base_dir=str(tmp_path))


def test_imported_scene_fixture_matches_reference_file_when_available(cpbrt):
    ref = "/root/reference/scenes/cornel_box.pbrt"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present (GPU box)")
    scene_a, scene_b = cpbrt.Scene.load(ref), cpbrt.Scene.load(os.path.join(SCENES, "cornel_box.pbrt"))   # owners of the arrays below
    a, b = scene_a.desc, scene_b.desc
    assert a.n_primitives == b.n_primitives and a.n_triangles == b.n_triangles
    ta = np.ctypeslib.as_array(a.triangles, shape=(int(a.n_triangles),)).view(np.uint8)
    tb = np.ctypeslib.as_array(b.triangles, shape=(int(b.n_triangles),)).view(np.uint8)
    assert np.array_equal(ta, tb)
    assert list(a.lights[0].L) == list(b.lights[0].L)


def test_exr_texture_and_environment_map(cpbrt, tmp_path):
    """OpenEXR images (texture/image/exr.rs:7-61): channels R, G, B as f16 / f32 / u32, NONE / RLE / ZIPS /
    ZIP scanline blocks, data window not at the origin; as an `imagemap` texture (with the 1.05 inverse
    gamma of spectrum textures) and as the `mapname` of an infinite light (raw values)."""
    rng = np.random.RandomState(5)
    H, W = 37, 21                                     # 37 rows: ZIP blocks of 16 + a ragged last block of 5
    img = (rng.rand(H, W, 3) * 4).astype(np.float32)
    img[5:9, :, :] = 0.25                             # constant rows: runs for RLE, blocks that really shrink
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [8] "integer yresolution" [8]\nIntegrator "path"\nWorldBegin\n')
    for comp in (0, 1, 2, 3):
        for ptype in (1, 2, 0):
            name = "env_%d_%d.exr" % (comp, ptype)
            src = np.floor(img * 50) if ptype == 0 else img
            write_exr(str(tmp_path / name), src, comp, ptype, origin=(3, -2))
            expect = src.astype(np.float16).astype(np.float32) if ptype == 1 else src.astype(np.float32)
            s = cpbrt.Scene.parse(head + 'LightSource "infinite" "string mapname" "%s"\nMaterial "matte"\nShape "sphere"\nWorldEnd\n' % name, base_dir=str(tmp_path))
            d = s.desc
            tex = d.textures[d.lights[0].tex]
            assert (tex.kind, tex.channels, tex.width, tex.height) == (cpbrt.abi.TEX_IMAGE, 3, W, H)
            got = np.ctypeslib.as_array(d.texels, shape=(int(d.n_texels),))[tex.offset:tex.offset + W * H * 3].reshape(H, W, 3)
            assert np.array_equal(got, expect), (comp, ptype)
    # spectrum imagemap: same loader, then the inverse gamma of texture_map.rs:41-49
    write_exr(str(tmp_path / "tex.exr"), img / 4, 3, 1)
    s = cpbrt.Scene.parse(head + 'Texture "t" "spectrum" "imagemap" "string filename" "tex.exr"\n'
                          'Material "matte" "texture Kd" "t"\nShape "sphere"\nWorldEnd\n', base_dir=str(tmp_path))
    d = s.desc
    tex = d.textures[d.materials[0].tex[0]]
    v = (img / 4).astype(np.float16).astype(np.float32)
    expect = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.05) ** 2.4)
    got = np.ctypeslib.as_array(d.texels, shape=(int(d.n_texels),))[tex.offset:tex.offset + W * H * 3].reshape(H, W, 3)
    np.testing.assert_allclose(got, expect, rtol=2e-5, atol=1e-7)
    # what the reference panics on is an error here (extra channels); lossy codecs are rejected by name
    bad = bytearray(open(str(tmp_path / "tex.exr"), "rb").read())
    i = bad.index(b"compression\0compression\0") + len(b"compression\0compression\0") + 4
    bad[i] = 5                                        # PXR24 (lossy)
    open(str(tmp_path / "pxr24.exr"), "wb").write(bytes(bad))
    with pytest.raises(cpbrt.CpbrtError, match="PXR24"):
        cpbrt.Scene.parse(head + 'LightSource "infinite" "string mapname" "pxr24.exr"\nMaterial "matte"\nShape "sphere"\nWorldEnd\n', base_dir=str(tmp_path))
    bad = bytearray(open(str(tmp_path / "tex.exr"), "rb").read())
    i = bad.index(b"chlist\0") + len(b"chlist\0") + 4
    bad[i] = ord("A")                                 # channel B renamed to A
    open(str(tmp_path / "alpha.exr"), "wb").write(bytes(bad))
    with pytest.raises(cpbrt.CpbrtError, match="neither R, G nor B"):
        cpbrt.Scene.parse(head + 'LightSource "infinite" "string mapname" "alpha.exr"\nMaterial "matte"\nShape "sphere"\nWorldEnd\n', base_dir=str(tmp_path))


@pytest.mark.parametrize("name", ["piz_half", "piz_float", "zip_half", "zips_float", "rle_half", "none_float"])
def test_exr_files_written_by_openexr(cpbrt, name):
    """Fixtures encoded by the OpenEXR library itself (through OpenCV, tests/golden/make_exr_golden.py), so
    the reader is checked against an encoder that shares nothing with this repository -- including PIZ
    (Huffman + wavelet), which the test-suite's own writer does not produce."""
    src = np.load(os.path.join(GOLDEN, "exr", "exr_source.npz"))["rgb"]
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [8] "integer yresolution" [8]\nIntegrator "path"\nWorldBegin\n')
    s = cpbrt.Scene.parse(head + 'LightSource "infinite" "string mapname" "%s.exr"\nMaterial "matte"\nShape "sphere"\nWorldEnd\n' % name,
                          base_dir=os.path.join(GOLDEN, "exr"))
    d = s.desc
    tex = d.textures[d.lights[0].tex]
    H, W, _ = src.shape
    assert (tex.width, tex.height, tex.channels) == (W, H, 3)
    got = np.ctypeslib.as_array(d.texels, shape=(int(d.n_texels),))[tex.offset:tex.offset + W * H * 3].reshape(H, W, 3)
    expect = src.astype(np.float16).astype(np.float32) if name.endswith("half") else src
    assert np.array_equal(got, expect)


def test_lexer_number_fast_path_equals_strtof(cpbrt, tmp_path):
    """The scene lexer reads the common numeric literals without strtof (csrc/host/parser.cpp, parse_number: integer
    grammar, or one exact double operation + a midpoint guard, strtof for everything else).  tests/lexer_number_check.cpp
    compares it with strtof / the i32 grammar on 600 000 literals of every shape a scene file holds -- "%.9g" and
    "%.17g" floats over the whole f32 range, indices, exponents, 20-digit mantissas, exact midpoints of two floats,
    negative zeros, leading zeros -- bit for bit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "curry-pbrt_b200")
    exe = str(tmp_path / "lexer_number_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(root, "tests", "lexer_number_check.cpp"), "-o", exe,
                           "-L", pkg, "-l:libcpbrt.so", "-Wl,-rpath," + pkg, "-lm"])
    out = subprocess.run([exe, "600000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout[-2000:]


def test_cornell_triangles_against_an_independent_reader(cpbrt):
    """The host parser and scene builder are shared by the product and by the oracle's input, so an error there would
    cancel in every parity test.  This reads the triangles of scenes/cornel_box.pbrt with a reader that shares nothing
    with csrc/host (a regular expression over the text + numpy): every `trianglemesh` block in file order, one triangle
    per index triple with the vertices in index order (triangle.rs:30-36) -- the object "ball_and_box" is instanced
    right after its definition, so file order is push order (scene.rs:84-140) -- and compares them with the flattened
    scene description vertex for vertex, bit for bit; the 32 + 1 primitives must all survive the camera clipper."""
    import re
    text = open(os.path.join(SCENES, "cornel_box.pbrt")).read()
    text = re.sub(r"#[^\n]*", "", text)
    want = []
    for m in re.finditer(r'Shape\s+"trianglemesh"\s+"integer indices"\s*\[([^\]]*)\]\s*"point P"\s*\[([^\]]*)\]', text):
        idx = np.array(m.group(1).split(), np.int64).reshape(-1, 3)
        pts = np.array(m.group(2).split(), np.float32).reshape(-1, 3)
        want.append(pts[idx])                       # [n_tris, 3 vertices, xyz]
    want = np.concatenate(want)
    assert want.shape == (32, 3, 3)
    s = cpbrt.Scene.load(os.path.join(SCENES, "cornel_box.pbrt"))
    d = s.desc
    tri_prims = [d.primitives[i] for i in range(d.n_primitives) if d.primitives[i].shape.kind == cpbrt.abi.SHAPE_TRIANGLE]
    got = np.array([[[d.triangles[p.shape.data].p[v][c] for c in range(3)] for v in range(3)] for p in tri_prims], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert all(p.shape.xform == -1 for p in tri_prims)            # no transform is in force where the meshes are declared
    # the first two triangles are the light quad: emissive primitives, one AreaLight each (scene.rs:91-95)
    assert [p.light for p in tri_prims[:2]] == [0, 1] and all(p.light == -1 for p in tri_prims[2:])


def test_big_mesh_parallel_clipper_equals_sequential(cpbrt):
    """Meshes of >= 65 536 triangles run the camera clipper (perspective.rs:54-78) on all host threads and create
    their triangles in parallel (csrc/host/scene_build.cpp); smaller ones take the sequential path.  The same 80 000
    triangles -- a strip far wider than the view, so that many are clipped -- as ONE mesh and as four meshes of
    20 000 must give the same primitives in the same order."""
    rng = np.random.default_rng(5)
    n = 80000
    c = np.stack([rng.uniform(-40, 40, n), rng.uniform(-3, 3, n), rng.uniform(-2, 30, n)], 1).astype(np.float32)
    tris = (c[:, None, :] + rng.normal(scale=0.05, size=(n, 3, 3))).astype(np.float32)
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [64] "integer yresolution" [48]\nIntegrator "path"\nWorldBegin\nMaterial "matte"\n')

    def mesh(t):
        return 'Shape "trianglemesh" "integer indices" [%s] "point P" [%s]\n' % (
            " ".join(str(i) for i in range(3 * len(t))), " ".join(repr(float(v)) for v in t.ravel()))
    one = cpbrt.Scene.parse(head + mesh(tris) + "WorldEnd\n")
    four = cpbrt.Scene.parse(head + "".join(mesh(tris[k:k + 20000]) for k in range(0, n, 20000)) + "WorldEnd\n")
    a, b = one.desc, four.desc
    assert a.n_primitives == b.n_primitives and 0 < a.n_primitives < n          # the clipper dropped some, kept some
    assert a.n_triangles == b.n_triangles == n

    def kept(d):
        idx = np.array([d.primitives[i].shape.data for i in range(d.n_primitives)])
        pts = np.ctypeslib.as_array(C.cast(d.triangles, C.POINTER(C.c_float)), shape=(d.n_triangles, 15))
        return idx, pts[idx, :9].copy()
    import ctypes as C
    ia, pa = kept(a); ib, pb = kept(b)
    assert np.array_equal(ia, ib) and np.array_equal(pa.view(np.uint32), pb.view(np.uint32))


def test_clipper_against_an_independent_restatement(cpbrt):
    """The camera clipper is host code shared by the product and the oracle, so no parity test sees it.  Restated here
    from the Rust alone (camera/mod.rs:58-63, camera/perspective.rs:54-77): a material primitive goes when BOTH its
    camera-space bound's min and max points (`bound[corner]` indexes modulo 2 -- only those two of the eight corners are
    ever looked at) project behind the camera or off the film; emitters always stay.  A height field seen from close by
    loses half its triangles; which ones must agree for every triangle whose decision does not hinge on rounding
    (the world-to-camera matrix is re-derived here in float64, so decisions within 1e-3 pixel are set aside)."""
    from test_oracle import _synthetic_mesh_scene
    nx, nz = 40, 30
    text = _synthetic_mesh_scene(nx, nz, seed=5, res=64).replace("LookAt 278 273 -800 278 273 0 0 1 0", "LookAt 278 120 -100 278 20 300 0 1 0")
    scene = cpbrt.Scene.parse(text)
    d = scene.desc
    import re
    m = re.search(r'Material "matte"[^\n]*\nShape "trianglemesh" "integer indices" \[([^\]]*)\] "point P" \[([^\]]*)\]', text)
    idx = np.array(m.group(1).split(), np.int64).reshape(-1, 3)
    P = np.array([np.float32(float(v)) for v in m.group(2).split()], np.float32).reshape(-1, 3)
    tris = P[idx]                                                   # every triangle of the mesh, file order
    assert len(tris) == 2 * nx * nz
    world_to_camera = np.linalg.inv(np.array(list(d.camera.camera_to_world), np.float64).reshape(4, 4))
    camera_to_raster = np.array(list(d.camera.camera_to_raster), np.float64).reshape(4, 4)
    width, height = float(d.film_width), float(d.film_height)

    def decide(tri):
        lo, hi = tri.min(0).astype(np.float64), tri.max(0).astype(np.float64)
        corners = np.array([[(lo, hi)[(i >> c) & 1][c] for c in range(3)] + [1.0] for i in range(8)])
        cam = corners @ world_to_camera.T
        cam = cam[:, :3] / cam[:, 3:4]
        keep, margin = False, np.inf
        for p in (cam.min(0), cam.max(0)):
            h = camera_to_raster @ np.append(p, 1.0)
            x, y, z = h[:3] / h[3]
            margin = min(margin, abs(z), abs(x), abs(x - width), abs(y), abs(y - height))
            if not (z < 0 or x >= width or x < 0 or y >= height or y < 0):
                keep = True
        return keep, margin

    kept_actual = set()
    for i in range(int(d.n_primitives)):
        if d.primitives[i].material >= 0:
            kept_actual.add(np.array([[d.triangles[d.primitives[i].shape.data].p[k][c] for c in range(3)] for k in range(3)], np.float32).tobytes())
    assert int(d.n_primitives) - len(kept_actual) == 2 and len({t.tobytes() for t in tris}) == len(tris)
    robust = kept = 0
    for tri in tris:
        keep, margin = decide(tri)
        if margin > 1e-3:
            robust += 1
            kept += keep
            assert (tri.tobytes() in kept_actual) == keep
    assert robust > 0.99 * len(tris) and 0.2 * len(tris) < kept < 0.8 * len(tris)
