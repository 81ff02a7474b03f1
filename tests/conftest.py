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
