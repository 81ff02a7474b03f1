"""Whole-image comparison of the C++ oracle with the independent numpy restatement
(oracle/{sah,trace,path}_numpy.py) on scenes/cornel_box.pbrt.  CPU only, slow (about 400 paths/s).

    python tools/independent_restatement_check.py WIDTH HEIGHT SPP [MAXDEPTH] > profiles/...json
"""
import importlib
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
cpbrt = importlib.import_module("curry-pbrt_b200")
import oracle_api  # noqa: E402
from conftest import SCENES, scene_text  # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    width, height, spp = (int(v) for v in sys.argv[1:4])
    depth = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    sah, trace, path = load("sah_numpy"), load("trace_numpy"), load("path_numpy")
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=width, yres=height, spp=spp), base_dir=SCENES)
    world = oracle_api.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    mine.max_depth = depth

    def halton_value(x, y, s):
        dims = np.arange(128, dtype=np.uint32)
        _, v = world.halton(np.full(128, x, np.uint32), np.full(128, y, np.uint32), np.full(128, s, np.uint32), dims)
        return lambda dim: v[dim]

    def camera_ray_at(fx, fy):
        r = world.camera_ray_at(fx, fy)
        return r[0:3], r[4:7]

    t0 = time.time()
    film = path.render(mine, width, height, spp, halton_value, camera_ray_at)
    seconds = time.time() - t0
    ref, _, st = world.render(max_depth=depth)
    same = (film.view(np.uint32) == ref.view(np.uint32)).all(axis=2)
    print(json.dumps({
        "scene": "scenes/cornel_box.pbrt", "film": [width, height], "spp": spp, "maxdepth": depth,
        "paths": width * height * spp, "numpy_restatement_seconds": round(seconds, 1),
        "bvh_order_equal": bool(np.array_equal(mine.search.tree.order, world.bvh_order())),
        "rays_closest": [mine.rays_closest, st["rays_closest"]], "rays_any": [mine.rays_any, st["rays_any"]],
        "pixels": int(same.size), "pixels_bit_equal": int(same.sum()),
        "max_abs_difference": float(np.abs(film - ref).max()), "image_mean": float(ref.mean())}))


if __name__ == "__main__":
    main()
