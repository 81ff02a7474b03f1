#!/usr/bin/env python3
"""Exact SAH build time (cpbrt_build_bvh) on the synthetic height fields: host only, the device phases with the
pieces stitched on the host (CPBRT_BVH_RESIDENT=0), and the device-resident build (bounds to traversal layout on the
GPU, only the primitive order comes back).  One JSON line per size; all builds must give the same primitive order
and the resident and host layouts must describe the same tree.

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
            resident = r.bvh_device_layout()["resident"] if int(scene.desc.n_primitives) <= 2000000 else None
            row["gpu_minlen_%s" % ml] = dict(seconds=dt, device_levels=r.bvh_gpu_levels(), host_segments=r.bvh_host_segments(), resident=resident,
                                             same_order=bool(np.array_equal(order, ref)))
            os.environ["CPBRT_BVH_RESIDENT"] = "0"
            r.build_bvh()
            t0 = time.perf_counter(); order = r.build_bvh(); dt = time.perf_counter() - t0
            row["gpu_minlen_%s_hybrid" % ml] = dict(seconds=dt, device_levels=r.bvh_gpu_levels(), host_segments=r.bvh_host_segments(),
                                                    same_order=bool(np.array_equal(order, ref)))
            os.environ.pop("CPBRT_BVH_RESIDENT")
        os.environ.pop("CPBRT_BVH_BUILD", None); os.environ.pop("CPBRT_BVH_GPU_MINLEN", None)
        print(json.dumps(row), flush=True)
        r.close()


if __name__ == "__main__":
    main()
