#!/bin/bash
# round 2, GPU session f: new parity tests at config sizes + adversarial geometry, replay timing, bench
set -x
OUT=gpurun_out
TAG=r02_f
python -m pytest tests/test_gpu_config_sizes.py -x -q -s --durations=15 > $OUT/${TAG}_pytest_sizes.log 2>&1; echo "pytest sizes rc=$?" >> $OUT/${TAG}_pytest_sizes.log
tail -30 $OUT/${TAG}_pytest_sizes.log
python -m pytest tests/test_gpu_parity.py -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
python - <<'PY' 2>&1 | tee $OUT/${TAG}_replay.txt
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__
cpbrt = __graft_entry__.build()
scene = cpbrt.Scene.load('scenes/cornel_box.pbrt'); scene.set_spp(1024)
r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
import torch
film = torch.zeros(900*900*3, dtype=torch.float32, device='cuda')
for rng in ((0, 125), (899, 1024), (0, 125), (899, 1024), (512, 640)):
    _, st = r.render(sample_begin=rng[0], sample_end=rng[1], device_film_ptr=film.data_ptr())
    print('samples', rng, 'ms_total %.2f' % st.ms_total, 'launches', st.kernel_launches)
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-900 $OUT/${TAG}_bench.json
