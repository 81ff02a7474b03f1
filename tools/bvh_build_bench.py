#!/usr/bin/env python3
"""Exact SAH build time, host-only vs device top levels + host finish (cpbrt_build_bvh), on the
synthetic height fields.  One JSON line per size; both builds must give the same primitive order.

    python tools/bvh_build_bench.py [--sizes 1000x500,2500x2000] [--min-lens 4096]
"""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000x500,2500x2000")
    ap.add_argument("--min-lens", default="4096")
    args = ap.parse_args()
    import numpy as np
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    for size in args.sizes.split(","):
        nx, nz = (int(v) for v in size.split("x"))
        scene = syn.load_heightfield(cpbrt, nx, nz, 64, 64, 1)
        r = cpbrt.Renderer(0); r.upload(scene)
        row = dict(size=size, primitives=int(scene.desc.n_primitives), host_threads=os.cpu_count())
        os.environ["CPBRT_BVH_BUILD"] = "host"
        t0 = time.perf_counter(); ref = r.build_bvh(); row["host_s"] = time.perf_counter() - t0
        os.environ["CPBRT_BVH_BUILD"] = "gpu"
        for ml in args.min_lens.split(","):
            os.environ["CPBRT_BVH_GPU_MINLEN"] = ml
            r.build_bvh()          # warm-up (CUB first-use, allocations)
            t0 = time.perf_counter(); order = r.build_bvh(); dt = time.perf_counter() - t0
            row["gpu_minlen_%s" % ml] = dict(seconds=dt, device_levels=r.bvh_gpu_levels(), same_order=bool(np.array_equal(order, ref)))
        os.environ.pop("CPBRT_BVH_BUILD", None); os.environ.pop("CPBRT_BVH_GPU_MINLEN", None)
        print(json.dumps(row), flush=True)
        r.close()


if __name__ == "__main__":
    main()
