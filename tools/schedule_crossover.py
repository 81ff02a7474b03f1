#!/usr/bin/env python3
"""Loop- vs vote-scheduled traversal (bvh_trace.cuh) across tree sizes: random closest-hit rays,
shadow rays and a small full render per size, both schedules (CPBRT_SCHEDULE).  One JSON line per size.

    python tools/schedule_crossover.py [--sizes 16x8,64x32,200x100,1000x500]
"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16x8,64x32,200x100,1000x500")
    ap.add_argument("--rays", type=int, default=21)
    args = ap.parse_args()
    import numpy as np, torch
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    from conftest import lcg_rays
    for size in args.sizes.split(","):
        nx, nz = (int(v) for v in size.split("x"))
        scene = syn.load_heightfield(cpbrt, nx, nz, 960, 540, 16)
        r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
        rnd = lcg_rays(1 << args.rays, (0, 0, 0), (556, 548, 559), seed=7)
        sh = lcg_rays(1 << args.rays, (0, 0, 0), (556, 548, 559), seed=13)
        sh[:, 3] = np.float32(1 - 1e-5); sh[:, 4:7] *= np.float32(300)
        d_rnd, d_sh = torch.from_numpy(rnd).cuda(), torch.from_numpy(sh).cuda()
        out_c = torch.zeros(rnd.shape[0] * 16, dtype=torch.uint8, device="cuda")
        out_a = torch.zeros(sh.shape[0], dtype=torch.uint8, device="cuda")
        row = dict(size=size, primitives=int(scene.desc.n_primitives))
        ref = {}
        for sched in ("loop", "vote"):
            os.environ["CPBRT_SCHEDULE"] = sched
            ms_c = min(r.trace_closest_device(d_rnd.data_ptr(), rnd.shape[0], out_c.data_ptr()).ms_trace for _ in range(5))
            ms_a = min(r.trace_any_device(d_sh.data_ptr(), sh.shape[0], out_a.data_ptr()).ms_trace for _ in range(5))
            r.render()
            film, st = r.render(profile_stages=True)
            row[sched] = dict(closest_ms=ms_c, any_ms=ms_a, render_ms=st.ms_total, trace_closest_ms=st.ms_trace_closest, trace_shadow_ms=st.ms_trace_shadow)
            res = (out_c.cpu().numpy().copy(), out_a.cpu().numpy().copy(), film.copy())
            if ref:
                row["identical"] = all(np.array_equal(a, b) for a, b in zip(ref["r"], res))
            ref["r"] = res
        os.environ.pop("CPBRT_SCHEDULE", None)
        print(json.dumps(row), flush=True)
        r.close()


if __name__ == "__main__":
    main()
