#!/usr/bin/env python3
"""Traversal roofline on the BVH-traversal-bound configuration (C4: 1,000,008 triangles).

Ray batches resident in HBM -> cpbrt_trace_closest / cpbrt_trace_any -> kernel time from CUDA
events on the library's stream.  Bytes per ray = the shipped kernel's OWN work, counted by its
counting variant on the same rays (CPBRT_TRACE_COUNT_OWN):
    32 (ray in) + 16 (hit out; 1 for any-hit) + 64 * node steps + 48 * primitive tests
against this GPU's L2 read peak measured in the same run.  The reference's own visit counts
(far-first order, loose boxes; device counters cross-checked against the CPU oracle on
a subset) are printed next to it as the work the two-phase traversal avoids.  One JSON object per batch.

    python tools/traversal_roofline.py [--config C4] [--rays 22] [--oracle-rays 18] [--repeat 5]
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
    ap.add_argument("--config", default="C4")
    ap.add_argument("--rays", type=int, default=22, help="log2 of the random batch size")
    ap.add_argument("--oracle-rays", type=int, default=17, help="log2 of the subset the CPU oracle counts")
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    from conftest import lcg_rays
    cfg = dict(syn.CONFIGS[args.config]); cfg["spp"] = 1
    scene = syn.load_heightfield(cpbrt, **cfg)
    r = cpbrt.Renderer(0)
    t0 = time.time(); r.upload(scene); r.build_bvh(); build_s = time.time() - t0
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # measured read peaks of this GPU: 64 MB stays in the 126 MB L2, 4 GB streams from HBM
    l2_gbs = r.read_bandwidth(64 << 20, 50)
    hbm_read_gbs = r.read_bandwidth(4 << 30, 5)
    print(json.dumps(dict(probe="read_bandwidth", l2_64MB_gbs=l2_gbs, hbm_4GB_gbs=hbm_read_gbs, hbm_copy_gbs_measured_peaks=peak)))
    world = None
    if not args.no_oracle:
        import oracle_api
        world = oracle_api.OracleWorld(scene)

    batches = {
        "camera_rays_sample0": r.camera_rays(0),
        "random_rays_lcg7": lcg_rays(1 << args.rays, (0, 0, 0), (556, 548, 559), seed=7),
    }
    shadow = lcg_rays(1 << args.rays, (0, 0, 0), (556, 548, 559), seed=13)
    shadow[:, 3] = np.float32(1 - 1e-5); shadow[:, 4:7] *= np.float32(300)
    for name, rays in list(batches.items()) + [("shadow_rays_lcg13", shadow)]:
        any_hit = name.startswith("shadow")
        n = rays.shape[0]
        d_rays = torch.from_numpy(rays).cuda()
        d_out = torch.zeros(n * (1 if any_hit else 16), dtype=torch.uint8, device="cuda")
        fn = r.trace_any_device if any_hit else r.trace_closest_device
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ms = []
        for i in range(args.repeat + 3):
            flush.fill_(i & 0xFF)            # evict L2 between timed launches
            torch.cuda.synchronize()
            st = fn(d_rays.data_ptr(), n, d_out.data_ptr())
            if i >= 3:
                ms.append(st.ms_trace)
        dev = fn(d_rays.data_ptr(), n, d_out.data_ptr(), count=True)          # reference-order traversal with its visit counters
        own = fn(d_rays.data_ptr(), n, d_out.data_ptr(), count="own")        # the shipped traversal counting its own work
        io = 32 + (1 if any_hit else 16)
        ref_per_ray = io + (32.0 * dev.trace_parent_visits + 48.0 * dev.trace_prim_tests) / n
        own_per_ray = io + (64.0 * own.trace_parent_visits + 48.0 * own.trace_prim_tests) / n
        out = dict(config=args.config, batch=name, kind="any" if any_hit else "closest", rays=n, primitives=int(scene.desc.n_primitives),
                   ms_avg=float(np.mean(ms)), ms_min=float(np.min(ms)), mrays_s=n / (np.mean(ms) * 1e-3) / 1e6,
                   own_node_steps_per_ray=own.trace_parent_visits / n, own_prim_tests_per_ray=own.trace_prim_tests / n, own_bytes_per_ray=own_per_ray,
                   replays=int(own.trace_slow_paths),
                   reference_parent_visits_per_ray=dev.trace_parent_visits / n, reference_prim_tests_per_ray=dev.trace_prim_tests / n,
                   reference_bytes_per_ray=ref_per_ray, work_avoided_vs_reference=ref_per_ray / own_per_ray,
                   bvh_build_upload_s=build_s, l2="flushed with a 256 MB write before every timed launch")
        if world is not None:
            m = min(n, 1 << args.oracle_rays)
            sub = rays[:: max(1, n // m)][:m]
            if any_hit:
                res, cnt = world.trace_any(sub, count=True, threads=os.cpu_count())
            else:
                res, cnt = world.trace_closest(sub, count=True, threads=os.cpu_count())
            out.update(oracle_parent_visits_per_ray=int(cnt[0]) / len(sub), oracle_prim_tests_per_ray=int(cnt[1]) / len(sub), oracle_subset=len(sub))
        achieved = own_per_ray * n / (np.mean(ms) * 1e-3) / 1e9
        out.update(roofline=dict(bound="l2", achieved=achieved, peak=l2_gbs, unit="GB/s", frac=achieved / l2_gbs,
                                 note="own work: node steps x 64 B + primitive tests x 48 B + ray in / result out; nodes + triangles of C4 (94 MB) fit the L2; "
                                      "peak = read-only kernel over 64 MB in this run"),
                   hbm_fraction_of_the_same_bytes=achieved / peak)
        print(json.dumps(out))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
