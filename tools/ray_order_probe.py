#!/usr/bin/env python3
"""How much would the vote-scheduled traversal gain from rays that arrive sorted?  C4 (1,000,008 triangles), the 2^22
random rays of the roofline batch traced as they come, grouped by direction octant, and grouped by octant + origin
cell (16^3 grid).  Upper bound for what binning the closest-hit queue when it is written could buy."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np, torch
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    from conftest import lcg_rays
    cfg = dict(syn.CONFIGS["C4"]); cfg["spp"] = 1
    scene = syn.load_heightfield(cpbrt, **cfg)
    r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
    n = 1 << 22
    rays = lcg_rays(n, (0, 0, 0), (556, 548, 559), seed=7)
    octant = (rays[:, 4] < 0).astype(np.int64) | ((rays[:, 5] < 0).astype(np.int64) << 1) | ((rays[:, 6] < 0).astype(np.int64) << 2)
    cell = np.clip((rays[:, 0:3] / np.array([556, 548, 559], np.float32) * 16).astype(np.int64), 0, 15)
    cell_id = (cell[:, 0] * 16 + cell[:, 1]) * 16 + cell[:, 2]
    orders = {"as_generated": np.arange(n), "by_octant": np.argsort(octant, kind="stable"),
              "by_octant_and_origin_cell": np.argsort(octant * 4096 + cell_id, kind="stable"),
              "by_origin_cell": np.argsort(cell_id, kind="stable")}
    hits = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, order in orders.items():
        d = torch.from_numpy(np.ascontiguousarray(rays[order])).cuda()
        ms = []
        for i in range(5):
            flush.fill_(i); torch.cuda.synchronize()
            st = r.trace_closest_device(d.data_ptr(), n, hits.data_ptr())
            if i:
                ms.append(st.ms_trace)
        out[name] = dict(ms=float(np.mean(ms)), grays_s=n / np.mean(ms) / 1e6)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
