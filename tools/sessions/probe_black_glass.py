"""One-off probe (GPU box): the Cornell box with a black glass ball (NaN samples, see
tests/test_oracle.py::test_nan_samples_against_independent_numpy_restatement) -- device against the C++ oracle."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
cpbrt = importlib.import_module("curry-pbrt_b200")
import oracle_api  # noqa: E402
from conftest import SCENES, scene_text  # noqa: E402

text = scene_text("cornel_box.pbrt", xres=20, yres=20, spp=3).replace('Material "glass"', 'Material "glass" "rgb Kr" [0 0 0] "rgb Kt" [0 0 0]')
scene = cpbrt.Scene.parse(text, base_dir=SCENES)
r = cpbrt.Renderer(0)
r.upload(scene)
r.build_bvh()
got, stats = r.render()
ref, _, st = oracle_api.OracleWorld(scene).render()
finite = ~np.isnan(ref).any(axis=2)
rel = (np.abs(got - ref) / (np.abs(ref) + 1e-3)).max(axis=2)
print(json.dumps({"device_nan_dropped": int(stats.nan_dropped), "oracle_nan_dropped": int(st["nan_dropped"]),
                  "nan_pixels_device": int(np.isnan(got).any(axis=2).sum()), "nan_pixels_oracle": int((~finite).sum()),
                  "nan_pixels_same": bool(np.array_equal(np.isnan(got).any(axis=2), ~finite)),
                  "finite_pixels_over_1e-3": int((rel[finite & ~np.isnan(got).any(axis=2)] > 1e-3).sum()),
                  "rays_any": [int(stats.rays_any), int(st["rays_any"])]}))
