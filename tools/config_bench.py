#!/usr/bin/env python3
"""Render the other BASELINE.json configurations on one GPU and print one JSON line each
(bench.py itself stays on the headline configuration, Cornell box at 1024 spp).

    C1  scenes/cornel_box.pbrt, 900x900, 16 spp (the reference's own CPU-runnable case)
    C2  scenes/sphere.pbrt, 1600x900, 1024 spp, directlighting
    C3  Cornell box with box-texture.png as matte Kd, 1920x1080, 1024 spp
    C4  1,000,008-triangle height field, 1920x1080, 256 spp
    C5  10,000,008-triangle height field, 3840x2160 (4096 spp is the 8-GPU job; --c5-spp samples here)

Each line carries Msamples/s and Mrays/s (device time of cpbrt_render, scene and BVH resident),
the upload and BVH-build seconds, and two size-independent checks done at full size: rendering
the first two samples as [0,1) + [1,2) equals [0,2), and a repeated render is bit-identical.
--cpu adds the oracle's rate on a bounded crop of the same configuration (all host threads).

    python tools/config_bench.py [--configs C2,C3,C4] [--c5-spp 16] [--cpu]
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4")
    ap.add_argument("--c5-spp", type=int, default=16)
    ap.add_argument("--spp-scale", type=float, default=1.0, help="scale every spp (quick runs)")
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    from conftest import SCENES, c3_text, scene_text
    syn = importlib.import_module("curry-pbrt_b200.synthetic")

    def make(name):
        if name == "C1":
            return cpbrt.Scene.parse(scene_text("cornel_box.pbrt", spp=16), base_dir=SCENES), "scenes/cornel_box.pbrt 900x900 path d5"
        if name == "C2":
            return cpbrt.Scene.parse(scene_text("sphere.pbrt", spp=1024), base_dir=SCENES), "scenes/sphere.pbrt 1600x900 directlighting d0"
        if name == "C3":
            return cpbrt.Scene.parse(c3_text(scene_text), base_dir=SCENES), "Cornell box + box-texture.png matte Kd, 1920x1080 path d5"
        cfg = dict(syn.CONFIGS[name])
        if name == "C5":
            cfg["spp"] = args.c5_spp
        syn.write_heightfield_files(syn.default_dir(), **cfg)      # the mesh file is an input: written before the parse is timed
        return syn.load_heightfield(cpbrt, **cfg), "height field %d x %d quads as .pbrt + binary .ply (plymesh), %dx%d path d5" % (cfg["nx"], cfg["nz"], cfg["xres"], cfg["yres"])

    for name in args.configs.split(","):
        t0 = time.perf_counter()
        scene, what = make(name)
        parse_s = time.perf_counter() - t0
        spp = max(2, int(round(scene.desc.sampler.spp * args.spp_scale)))
        scene.set_spp(spp)
        W, H = scene.resolution
        r = cpbrt.Renderer(0)
        t0 = time.perf_counter(); r.upload(scene); upload_s = time.perf_counter() - t0
        t0 = time.perf_counter(); r.build_bvh(); build_s = time.perf_counter() - t0
        # size-independent checks at full resolution
        a, _ = r.render(sample_begin=0, sample_end=1)
        b, _ = r.render(sample_begin=1, sample_end=2)
        ab, _ = r.render(sample_begin=0, sample_end=2)
        ab2, _ = r.render(sample_begin=0, sample_end=2)
        compose_ok = bool(np.allclose((a + b) / 2, ab, rtol=2e-5, atol=1e-6))
        repeat_ok = bool(np.array_equal(ab.view(np.uint32), ab2.view(np.uint32)))
        del a, b, ab, ab2
        r.render(sample_begin=0, sample_end=min(spp, 4))          # warm-up
        film, st = r.render(profile_stages=False)
        film, st = r.render(profile_stages=True)
        sec = st.ms_total * 1e-3
        out = dict(config=name, workload=what, film=[W, H], spp=spp, primitives=int(scene.desc.n_primitives),
                   msamples_s=W * H * spp / sec / 1e6, mrays_s=(st.rays_closest + st.rays_any) / sec / 1e6,
                   mrays_traced_s=(st.rays_closest_traced + st.rays_any) / sec / 1e6, render_s=sec,
                   stage_ms=dict(trace_closest=st.ms_trace_closest, trace_shadow=st.ms_trace_shadow, shade=st.ms_shade, raygen_film=st.ms_raygen_film),
                   parse_s=parse_s, upload_s=upload_s, bvh_build_s=build_s, nan_dropped=int(st.nan_dropped),
                   film_mean=float(film.mean()), checks=dict(sample_ranges_compose=compose_ok, repeat_bit_identical=repeat_ok))
        if args.cpu:
            import oracle_api
            cw, ch = max(16, W // 12), max(16, H // 12)
            rect = ((W - cw) // 2, (H - ch) // 2, (W - cw) // 2 + cw, (H - ch) // 2 + ch)
            cpu_spp = min(spp, 8)
            scene.set_spp(cpu_spp)
            world = oracle_api.OracleWorld(scene)
            _, csec, cst = world.render(rect=rect, threads=os.cpu_count() or 1)
            n = cw * ch * cpu_spp
            out["cpu_baseline"] = dict(value=n / csec / 1e6, unit="Msamples/s", cores=os.cpu_count() or 1, kind="port",
                                       sample="centre crop %s at %d spp (%d samples), BVH build excluded" % (list(rect), cpu_spp, n))
        print(json.dumps(out), flush=True)
        r.close()


if __name__ == "__main__":
    sys.exit(main())
