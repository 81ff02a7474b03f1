"""Whole-image or cropped comparison of the C++ oracle with the independent numpy restatement
(oracle/{front,sah,trace,path}_numpy.py: sampler, camera, BVH, search, integrator -- nothing of the C++
oracle on that side) on the BASELINE scenes at any film size.  CPU only, slow (about 400 paths/s).

    python tools/independent_restatement_check.py --scene cornell --film 96 96 --spp 4
    python tools/independent_restatement_check.py --scene cornell --film 900 900 --spp 16 --rect 435 435 465 465 --rect 0 0 20 20
"""
import argparse
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
from conftest import SCENES, c3_text, scene_text  # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def scene_file_text(which, width, height, spp):
    if which == "cornell":
        return scene_text("cornel_box.pbrt", xres=width, yres=height, spp=spp)
    if which == "sphere":
        return scene_text("sphere.pbrt", xres=width, yres=height, spp=spp)
    text = c3_text()
    for key, old, new in (("xresolution", 1920, width), ("yresolution", 1080, height), ("pixelsamples", 1024, spp)):
        text = text.replace('"integer %s" [%d]' % (key, old), '"integer %s" [%d]' % (key, new))
    return text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", choices=["cornell", "sphere", "textured"], default="cornell")
    ap.add_argument("--film", type=int, nargs=2, default=[96, 96])
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--rect", type=int, nargs=4, action="append", help="x0 y0 x1 y1 (repeatable); default the whole film")
    args = ap.parse_args()
    width, height = args.film
    spp = args.spp
    sah, trace, path, front = load("sah_numpy"), load("trace_numpy"), load("path_numpy"), load("front_numpy")
    scene = cpbrt.Scene.parse(scene_file_text(args.scene, width, height, spp), base_dir=SCENES)
    d = scene.desc
    world = oracle_api.OracleWorld(scene)
    mine = path.Scene(d, cpbrt.abi, sah, trace)
    integrator = "path" if int(d.integrator) == 0 else "directlighting"
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

    out = {"scene": args.scene, "film": [width, height], "spp": spp, "integrator": integrator, "maxdepth": int(d.max_depth),
           "bvh_order_equal": bool(np.array_equal(mine.search.tree.order, world.bvh_order())), "crops": []}
    for rect in (args.rect or [[0, 0, width, height]]):
        mine.rays_closest = mine.rays_any = 0
        t0 = time.time()
        film = path.render(mine, width, height, spp, halton_value, cam.generate_ray, rect=tuple(rect), integrator=integrator)
        seconds = time.time() - t0
        ref, _, st = world.render(rect=tuple(rect))
        x0, y0, x1, y1 = rect
        same = (film.view(np.uint32) == ref.view(np.uint32)).all(axis=2)[y0:y1, x0:x1]
        out["crops"].append({"rect": list(rect), "paths": int(same.size) * spp, "numpy_restatement_seconds": round(seconds, 1),
                             "rays_closest": [mine.rays_closest, st["rays_closest"]], "rays_any": [mine.rays_any, st["rays_any"]],
                             "pixels": int(same.size), "pixels_bit_equal": int(same.sum()),
                             "max_abs_difference": float(np.abs(film - ref).max()), "crop_mean": float(ref[y0:y1, x0:x1].mean())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
