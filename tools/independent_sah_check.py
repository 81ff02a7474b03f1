"""BVH order, node numbering and links of the C++ oracle against oracle/sah_numpy.py on a height-field mesh of nx x nz quads.

    python tools/independent_sah_check.py 400 250      # 200 000 triangles, about a minute
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import importlib, importlib.util
import numpy as np
cpbrt = importlib.import_module("curry-pbrt_b200")
import oracle_api
from test_oracle import _synthetic_mesh_scene
def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", name + ".py")); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
sah = load("sah_numpy")
nx, nz = int(sys.argv[1]), int(sys.argv[2])
sc = cpbrt.Scene.parse(_synthetic_mesh_scene(nx, nz, seed=2, res=16))
d = sc.desc; n = int(d.n_primitives)
T = np.ctypeslib.as_array(d.triangles, shape=(int(d.n_triangles),))
P = np.array([[list(d.triangles[d.primitives[i].shape.data].p[k]) for k in range(3)] for i in range(n)], np.float32)
t0=time.time(); tree = sah.build(*sah.triangle_bounds(P)); t1=time.time()
w = oracle_api.OracleWorld(sc)
ref = w.bvh_nodes()
import hashlib
print("sha256 of the numpy order as little-endian u32:", hashlib.sha256(np.asarray(tree.order, "<u4").tobytes()).hexdigest())
print(n, "numpy build s", round(t1-t0,1), "order equal", np.array_equal(tree.order, w.bvh_order()), "nodes", len(tree.a), len(ref["a"]),
      "links equal", np.array_equal(np.array(tree.a), ref["a"].astype(np.int64)) and np.array_equal(np.array(tree.b), ref["b"].astype(np.int64)))
