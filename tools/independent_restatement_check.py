"""Whole-image comparison of the C++ oracle with the independent numpy restatement
(oracle/{front,sah,trace,path}_numpy.py: sampler, camera, BVH, search, integrator -- nothing of the C++
oracle on that side) on scenes/cornel_box.pbrt.  CPU only, slow (about 400 paths/s).

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
    sah, trace, path, front = load("sah_numpy"), load("trace_numpy"), load("path_numpy"), load("front_numpy")
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=width, yres=height, spp=spp), base_dir=SCENES)
    world = oracle_api.OracleWorld(scene)
    mine = path.Scene(scene.desc, cpbrt.abi, sah, trace)
    mine.max_depth = depth

    d = scene.desc
    perm = np.ctypeslib.as_array(d.sampler.permutation, shape=(3682913,))
    cam = front.PerspectiveCamera(list(d.camera.raster_to_camera), list(d.camera.camera_to_world) if d.camera.has_camera_to_world else None)

    def halton_value(x, y, s):
        smp = front.HaltonSampler(spp, (width, height), perm)
        smp.set_pixel(x, y)
        for _ in range(s):
            smp.next_sample()

        def value(dim):
            smp.dim = dim
            return smp.get_sample()
        return value

    camera_ray_at = cam.generate_ray

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
