import sys, os, importlib
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import __graft_entry__
cp = __graft_entry__.build()
import oracle_api
from conftest import scene_text, SCENES
from test_gpu_parity import edge_case_rays
scene = cp.Scene.parse(scene_text("cornel_box.pbrt", xres=128, yres=128, spp=8), base_dir=SCENES)
r = cp.Renderer(0); r.upload(scene); order = r.build_bvh()
w = oracle_api.OracleWorld(scene)
rays = edge_case_rays()
g = r.trace_closest(rays); o = w.trace_closest(rays)
gc, _ = r.trace_closest(rays, count=True)
bad = np.nonzero((g["prim"] != o["prim"]) | (g["t"].view(np.uint32) != o["t"].view(np.uint32)))[0]
print("mismatches", len(bad), "of", len(rays), " counted-kernel mismatches", int((gc["prim"] != o["prim"]).sum()))
for i in bad[:25]:
    print(i, rays[i].tolist(), "gpu", g["prim"][i], g["t"][i], "oracle", o["prim"][i], o["t"][i])
