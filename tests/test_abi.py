"""The C-ABI library loads without a GPU, exports every symbol include/cpbrt.h declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(cpbrt):
    header = open(os.path.join(ROOT, "include", "cpbrt.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cpbrt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(cpbrt.lib, name), "libcpbrt.so does not export " + name
    assert declared == set(cpbrt.abi.ALL_SYMBOLS), "ctypes prototypes out of sync with the header"


def test_struct_sizes_match_header(cpbrt):
    abi = cpbrt.abi
    assert C.sizeof(abi.Ray) == 32 and C.sizeof(abi.Hit) == 16
    assert C.sizeof(abi.Transform) == 144
    assert C.sizeof(abi.Triangle) == 60
    assert C.sizeof(abi.Primitive) == 24


def _build_abi_smoke(tmp_path):
    """tests/abi_smoke.c as strict C99 against include/cpbrt.h, linked with libcpbrt.so."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    lib_dir = os.path.join(ROOT, "curry-pbrt_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe,
           "-L" + lib_dir, "-lcpbrt", "-lm", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_consumer_layout_matches_ctypes(cpbrt, tmp_path):
    """Every struct of the header as a C compiler lays it out (sizeof / offsetof printed by the C program)
    against the hand-written ctypes mirror: size, and offset + size of every field."""
    import subprocess
    abi = cpbrt.abi
    mirror = {"cpbrt_transform": abi.Transform, "cpbrt_triangle": abi.Triangle, "cpbrt_shape": abi.Shape, "cpbrt_primitive": abi.Primitive,
              "cpbrt_texture": abi.Texture, "cpbrt_material": abi.Material, "cpbrt_light": abi.Light, "cpbrt_camera": abi.Camera,
              "cpbrt_sampler": abi.Sampler, "cpbrt_scene_desc": abi.SceneDesc, "cpbrt_ray": abi.Ray, "cpbrt_hit": abi.Hit,
              "cpbrt_bvh_dump": abi.BvhDump, "cpbrt_render_params": abi.RenderParams, "cpbrt_stats": abi.Stats}
    out = subprocess.run([_build_abi_smoke(tmp_path), "--layout"], capture_output=True, text=True, check=True).stdout
    seen = {}
    for line in out.splitlines():
        st, field, off, size = line.split()
        cls = mirror[st]
        if field == ".":
            assert C.sizeof(cls) == int(size), st
            seen[st] = set()
            continue
        f = getattr(cls, field)
        assert (f.offset, f.size) == (int(off), int(size)), "%s.%s: C has (%s, %s), ctypes (%d, %d)" % (st, field, off, size, f.offset, f.size)
        seen[st].add(field)
    assert set(seen) == set(mirror)
    for st, cls in mirror.items():
        assert seen[st] == {name for name, _ in cls._fields_}, st
    # and every typedef'd struct of the header is covered
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "cpbrt.h")).read(), flags=re.S)
    assert set(re.findall(r"\}\s*(cpbrt_[a-z_]+)\s*;", header)) - {"cpbrt_status"} == set(mirror)


@pytest.mark.gpu
def test_c_consumer_end_to_end(cpbrt, tmp_path):
    """The same C program walks the boundary on the GPU: a hand-built scene through create / upload / build /
    trace / halton / render, checked against exact hits and the closed-form radiance of its centre pixel."""
    import subprocess
    r = subprocess.run([_build_abi_smoke(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout


def test_no_device_means_error_not_fallback(cpbrt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = cpbrt.lib.cpbrt_create(C.byref(ctx), 0)
    assert rc == cpbrt.abi.ERR_NO_DEVICE
    assert b"CUDA" in cpbrt.lib.cpbrt_last_error(None)
    # compute entry points reject a null context
    assert cpbrt.lib.cpbrt_render(None, None, None, None) == cpbrt.abi.ERR_BAD_CTX
    assert cpbrt.lib.cpbrt_trace_closest(None, None, 0, None, 0, None) == cpbrt.abi.ERR_BAD_CTX
    with pytest.raises(cpbrt.CpbrtError):
        cpbrt.Renderer(0)


def test_parse_errors_are_codes_not_crashes(cpbrt):
    with pytest.raises(cpbrt.CpbrtError) as e:
        cpbrt.Scene.load("/nonexistent/scene.pbrt")
    assert e.value.code == cpbrt.abi.ERR_PARSE
    with pytest.raises(cpbrt.CpbrtError):
        cpbrt.Scene.parse('Camera "perspective"\nWorldBegin\nWorldEnd\n')          # no Sampler
    with pytest.raises(cpbrt.CpbrtError):
        cpbrt.Scene.parse('Camera "perspective" "float fov" [40\n')                 # unpaired [ then EOF word


def test_hostile_scene_files_are_codes_not_crashes(cpbrt, tmp_path):
    """What overflows the stack in the reference (a file including itself, hundreds of thousands of nested
    blocks) and what would ask for terabytes (image headers that do not fit their file) must come back as an
    error code: the library lives inside the caller's process."""
    head = ('Camera "perspective" "float fov" [40]\nSampler "halton" "integer pixelsamples" [1]\n'
            'Film "image" "integer xresolution" [8] "integer yresolution" [8]\nIntegrator "path"\nWorldBegin\n')
    (tmp_path / "self.pbrt").write_text('Include "self.pbrt"\n')
    with pytest.raises(cpbrt.CpbrtError, match="Include nested"):
        cpbrt.Scene.parse(head + 'Include "self.pbrt"\nWorldEnd\n', base_dir=str(tmp_path))
    with pytest.raises(cpbrt.CpbrtError, match="nested deeper"):
        cpbrt.Scene.parse(head + "AttributeBegin\n" * 200000)
    # 400 levels are fine
    s = cpbrt.Scene.parse(head + 'LightSource "point"\nMaterial "matte"\n' + "AttributeBegin\n" * 400 + 'Translate 0 0 5\nShape "sphere"\n'
                          + "AttributeEnd\n" * 400 + "WorldEnd\n")
    assert s.desc.n_primitives == 1
    # a PNG whose header claims 60000 x 60000 pixels over a few bytes of data
    import struct
    import zlib
    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 60000, 60000, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    (tmp_path / "huge.png").write_bytes(png)
    with pytest.raises(cpbrt.CpbrtError, match="do not fit"):
        cpbrt.Scene.parse(head + 'Texture "t" "spectrum" "imagemap" "string filename" "huge.png"\nMaterial "matte" "texture Kd" "t"\nShape "sphere"\nWorldEnd\n',
                          base_dir=str(tmp_path))


def test_write_png_roundtrip(cpbrt, tmp_path):
    from PIL import Image
    film = np.zeros((4, 6, 3), np.float32)
    film[..., 0] = 0.5; film[1, 2] = (1.0, 0.0031308, 2.0); film[0, 0] = (-1.0, np.nan, 0.001)
    path = str(tmp_path / "x.png")
    cpbrt.write_png(path, film)
    img = np.asarray(Image.open(path))
    assert img.shape == (4, 6, 3)
    # gamma_correct(0.5) = 1.055*0.5^(1/2.4)-0.055 = 0.7354 -> 188
    assert img[2, 2, 0] == 188
    assert tuple(img[1, 2]) == (255, int(12.92 * 0.0031308 * 255 + 0.5), 255)
    assert tuple(img[0, 0]) == (0, 0, int(12.92 * 0.001 * 255 + 0.5))
