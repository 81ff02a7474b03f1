#!/bin/bash
# round 2, GPU session i: where do the extra milliseconds of the e2e path come from with two lanes?  + replay timing
OUT=gpurun_out
python - <<'PY' 2>&1 | tee $OUT/r02_i_e2e_probe.txt
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import __graft_entry__
cpbrt = __graft_entry__.build()
scene = cpbrt.Scene.load('scenes/cornel_box.pbrt'); scene.set_spp(1024)
r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
film_dev = torch.zeros(900*900*3, dtype=torch.float32, device='cuda')
pinned = torch.empty((900, 900, 3), dtype=torch.float32).pin_memory(); host = pinned.numpy()
pageable = np.zeros((900, 900, 3), np.float32)
def t(fn, n=4):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); st = fn(); torch.cuda.synchronize(); out.append((round(1e3*(time.perf_counter()-t0),1), round(st.ms_total,1)))
    return out
print('device film         ', t(lambda: r.render(device_film_ptr=film_dev.data_ptr())[1]))
print('pinned host film    ', t(lambda: r.render(film=host)[1]))
print('pageable host film  ', t(lambda: r.render(film=pageable)[1]))
def up_render():
    r.upload(scene); r.build_bvh(); return r.render(device_film_ptr=film_dev.data_ptr())[1]
print('upload+build+device ', t(up_render))
def up_render_host():
    r.upload(scene); r.build_bvh(); return r.render(film=host)[1]
print('upload+build+pinned ', t(up_render_host))
for rng in ((0, 1), (899, 900), (0, 1), (899, 900), (450, 451)):
    _, st = r.render(sample_begin=rng[0], sample_end=rng[1], device_film_ptr=film_dev.data_ptr())
    print('samples', rng, 'ms_total %.3f' % st.ms_total)
PY
