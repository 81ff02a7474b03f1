#!/usr/bin/env python3
"""Regenerates tests/golden/cornell_golden.npz from the CPU oracle.

The reference ships no golden vectors for this path and cannot be run here, so these
are ORACLE outputs: they pin the oracle against drift (and travel to the GPU box), they do not
pin it against the Rust binary.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__
cpbrt = __graft_entry__.build()
import oracle_api
from conftest import SCENES, scene_text

XRES = YRES = 48
SPP = 4
scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=XRES, yres=YRES, spp=SPP), base_dir=SCENES)
w = oracle_api.OracleWorld(scene)
rays = w.camera_rays(0)
hits = w.trace_closest(rays)
rng = np.random.RandomState(5)
n = 512
h_px = rng.randint(0, XRES, n).astype(np.uint32); h_py = rng.randint(0, YRES, n).astype(np.uint32)
h_sample = rng.randint(0, 4096, n).astype(np.uint32); h_dim = rng.randint(0, 64, n).astype(np.uint32)
h_index, h_value = w.halton(h_px, h_py, h_sample, h_dim)
film, _, st = w.render(threads=8)
np.savez_compressed(os.path.join(HERE, "cornell_golden.npz"), xres=XRES, yres=YRES, spp=SPP, bvh_order=w.bvh_order(),
                    camera_rays=rays, first_hit_prim=hits["prim"], first_hit_t=hits["t"],
                    h_px=h_px, h_py=h_py, h_sample=h_sample, h_dim=h_dim, h_index=h_index, h_value=h_value,
                    film=film, rays_closest=st["rays_closest"], rays_any=st["rays_any"])
print("wrote cornell_golden.npz", film.mean(axis=(0, 1)), st)
