"""Random small scenes (triangles, spheres bare and under Translate / Rotate / Scale, ObjectInstances, every material,
every light type, both integrators, thin lens or not) rendered by the C++ oracle and by the independent numpy
restatement; any pixel or ray count that differs is reported with the scene text.  CPU only.

    python tools/independent_restatement_fuzz.py --scenes 30 --seed 1
"""
import argparse
import importlib
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
cpbrt = importlib.import_module("curry-pbrt_b200")
import oracle_api  # noqa: E402
from conftest import SCENES  # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rgb(rng, lo=0.05, hi=0.9):
    if rng.rand() < 0.06:
        return "0 0 0"                 # black now and then: empty lobe lists, all-zero luminance tables, NaN samples
    return "%.3g %.3g %.3g" % tuple(rng.uniform(lo, hi, 3))


def material(rng, textured_ok):
    k = rng.randint(0, 9)
    if k == 0:
        return 'Material "matte" "rgb Kd" [%s]' % rgb(rng)
    if k == 1:
        return 'Material "matte" "rgb Kd" [%s] "float sigma" [%.3g]' % (rgb(rng), rng.uniform(1, 80))
    if k == 2:
        return 'Material "mirror" "rgb Kr" [%s]' % rgb(rng, 0.5, 1.0)
    if k == 3:
        return 'Material "glass" "rgb Kr" [%s] "rgb Kt" [%s] "float index" [%.3g]' % (rgb(rng, 0.3, 1), rgb(rng, 0.3, 1), rng.uniform(1.1, 2.2))
    if k == 4:
        return 'Material "plastic" "rgb Kd" [%s] "rgb Ks" [%s] "float roughness" [%.3g]' % (rgb(rng), rgb(rng), rng.uniform(0.01, 0.9))
    if k == 5:
        return ('Material "uber" "rgb Kd" [%s] "rgb Ks" [%s] "rgb Kr" [%s] "rgb Kt" [%s] "float uroughness" [%.3g] "float vroughness" [%.3g] '
                '"float index" [%.3g] "rgb opacity" [%s]' % (rgb(rng), rgb(rng), rgb(rng, 0, 0.4), rgb(rng, 0, 0.4), rng.uniform(0.02, 0.8),
                                                            rng.uniform(0.02, 0.8), rng.uniform(1.1, 2.0), rgb(rng, 0.4, 1.0)))
    if k == 6:
        return ('Material "translucent" "rgb Kd" [%s] "rgb Ks" [%s] "float roughness" [%.3g] "rgb reflect" [%s] "rgb transmit" [%s]'
                % (rgb(rng), rgb(rng), rng.uniform(0.02, 0.8), rgb(rng), rgb(rng)))
    if k == 7 and textured_ok:
        return 'Material "matte" "texture Kd" "boxtex"'
    return ('MakeNamedMaterial "a%d" "string type" "plastic" "rgb Kd" [%s] "rgb Ks" [%s] "float roughness" [%.3g]\n'
            'MakeNamedMaterial "b%d" "string type" "glass" "float index" [%.3g]\n'
            'Material "mix" "string namedmaterial1" "a%d" "string namedmaterial2" "b%d" "rgb amount" [%s]'
            % (k, rgb(rng), rgb(rng), rng.uniform(0.05, 0.5), k, rng.uniform(1.2, 1.8), k, k, rgb(rng, 0.1, 0.9)))


def transform(rng):
    parts = []
    if rng.rand() < 0.8:
        parts.append("Translate %.3g %.3g %.3g" % tuple(rng.uniform(-2, 2, 3) + np.array([0, 0, 6])))
    if rng.rand() < 0.5:
        parts.append("Rotate %.3g %.3g %.3g %.3g" % ((rng.uniform(-90, 90),) + tuple(rng.uniform(0.1, 1, 3))))
    if rng.rand() < 0.4:
        parts.append("Scale %.3g %.3g %.3g" % tuple(rng.uniform(0.5, 1.6, 3)))
    return "\n".join(parts) + ("\n" if parts else "")


def shape(rng):
    if rng.rand() < 0.5:
        return 'Shape "sphere" "float radius" [%.3g]' % rng.uniform(0.3, 1.2)
    p = rng.uniform(-1.5, 1.5, (4, 3))
    uv = ' "float uv" [0 0 1 0 1 1 0 1]' if rng.rand() < 0.5 else ""
    return 'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [%s]%s' % (" ".join("%.4g" % v for v in p.ravel()), uv)


def scene_text(rng, width, height, spp):
    integrator = "path" if rng.rand() < 0.6 else "directlighting"
    lens = ' "float lensradius" [%.3g] "float focaldistance" [%.3g]' % (rng.uniform(0.01, 0.2), rng.uniform(3, 9)) if rng.rand() < 0.3 else ""
    out = ['LookAt 0 %.3g -2 0 0 6 0 1 0' % rng.uniform(-0.5, 1.5), 'Camera "perspective"%s "float fov" [%.3g]' % (lens, rng.uniform(30, 70)),
           'Sampler "halton" "integer pixelsamples" [%d]' % spp,
           'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]' % (width, height),
           'Integrator "%s" "integer maxdepth" [%d]' % (integrator, rng.randint(1, 6)), "WorldBegin",
           'Texture "boxtex" "spectrum" "imagemap" "string filename" "box-texture.png"']
    if rng.rand() < 0.5:
        out.append('LightSource "infinite"' + (' "string mapname" "box-texture.png"' if rng.rand() < 0.5 else ""))
    if rng.rand() < 0.5:
        out.append('AttributeBegin\nTranslate %.3g %.3g %.3g\nLightSource "point" "rgb I" [%s]\nAttributeEnd' % (tuple(rng.uniform(-3, 3, 3) + np.array([0, 3, 4])) + (rgb(rng, 5, 30),)))
    if rng.rand() < 0.4:
        out.append('LightSource "distant" "rgb L" [%s] "point from" [0 0 0] "point to" [%.3g %.3g %.3g]' % ((rgb(rng, 0.2, 2),) + tuple(rng.uniform(-1, 1, 3) + np.array([0, -1.5, 0.5]))))
    for _ in range(rng.randint(1, 3)):            # emitters: one wrapper at most
        out.append('AttributeBegin\n%sAreaLightSource "diffuse" "rgb L" [%s]\n%s\nAttributeEnd' % (transform(rng), rgb(rng, 2, 12), shape(rng)))
    out.append('Material "matte" "rgb Kd" [0.6 0.6 0.6]\nShape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-6 -1.6 0 6 -1.6 0 6 -1.6 14 -6 -1.6 14]')
    for _ in range(rng.randint(2, 6)):
        out.append('AttributeBegin\n%s%s\n%s\nAttributeEnd' % (transform(rng), material(rng, True), shape(rng)))
    if rng.rand() < 0.5:
        emitter = '\nAreaLightSource "diffuse" "rgb L" [%s]\n%s' % (rgb(rng, 2, 12), shape(rng)) if rng.rand() < 0.5 else ""     # instanced emitter: the light keeps the un-instanced shape
        out.append('ObjectBegin "thing"\n%s\n%s\n%s%s\n%s%s\nObjectEnd' % (material(rng, True), shape(rng), transform(rng).replace(" 6\n", " 0\n"), material(rng, True), shape(rng), emitter))
        for _ in range(rng.randint(1, 3)):
            out.append('AttributeBegin\n%sObjectInstance "thing"\nAttributeEnd' % transform(rng))
    out.append("WorldEnd")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--film", type=int, nargs=2, default=[14, 12])
    ap.add_argument("--spp", type=int, default=2)
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    sah, trace, path, front = load("sah_numpy"), load("trace_numpy"), load("path_numpy"), load("front_numpy")
    width, height = args.film
    spp = args.spp
    summary = {"scenes": 0, "skipped": 0, "pixels": 0, "pixels_bit_equal": 0, "nan_dropped": 0, "mismatching_scenes": [], "rays": 0}
    for k in range(args.scenes):
        rng = np.random.RandomState(args.seed * 1000 + k)
        text = scene_text(rng, width, height, spp)
        try:
            scene = cpbrt.Scene.parse(text, base_dir=SCENES)
        except Exception as e:                     # e.g. every primitive clipped
            summary["skipped"] += 1
            continue
        d = scene.desc
        if int(d.n_primitives) == 0:
            summary["skipped"] += 1
            continue
        world = oracle_api.OracleWorld(scene)
        mine = path.Scene(d, cpbrt.abi, sah, trace)
        perm = np.ctypeslib.as_array(d.sampler.permutation, shape=(3682913,))
        cam = front.PerspectiveCamera(list(d.camera.raster_to_camera), list(d.camera.camera_to_world) if d.camera.has_camera_to_world else None,
                                      d.camera.lens_radius if d.camera.has_lens else None, d.camera.focal_distance if d.camera.has_lens else None)

        def halton_value(x, y, s):
            smp = front.HaltonSampler(spp, (width, height), perm)
            smp.set_pixel(x, y)
            for _ in range(s):
                smp.next_sample()

            def value(dim):
                smp.dim = dim
                return smp.get_sample()
            return value
        integrator = "path" if int(d.integrator) == 0 else "directlighting"
        film = path.render(mine, width, height, spp, halton_value, cam.generate_ray, integrator=integrator, camera_draws=bool(d.camera.has_lens))
        ref, _, st = world.render()
        same = ((film.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(film) & np.isnan(ref))).all(axis=2)
        ok = bool(same.all()) and (mine.rays_closest, mine.rays_any) == (st["rays_closest"], st["rays_any"]) and np.array_equal(mine.search.tree.order, world.bvh_order())
        summary["scenes"] += 1
        summary["pixels"] += int(same.size)
        summary["pixels_bit_equal"] += int(same.sum())
        summary["nan_dropped"] += int(st["nan_dropped"])
        summary["rays"] += int(st["rays_closest"]) + int(st["rays_any"])
        if not ok:
            summary["mismatching_scenes"].append(k)
            if args.dump:
                open(os.path.join(args.dump, "fuzz_%d_%d.pbrt" % (args.seed, k)), "w").write(text)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
