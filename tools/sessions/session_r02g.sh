#!/bin/bash
# round 2, GPU session g: two-lane pass pipelining A/B, replay kernel timing, C consumer, full bench line
set -x
OUT=gpurun_out
TAG=r02_g
python -m pytest tests/test_gpu_parity.py tests/test_abi.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
for lanes in 1 2; do
  CPBRT_LANES=$lanes python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_lanes$lanes.json 2> $OUT/${TAG}_bench.err; echo "bench lanes=$lanes rc=$?"
  python -c "
import json,sys
j=json.loads(open('$OUT/${TAG}_bench_lanes$lanes.json').read().strip().splitlines()[-1]); print('lanes $lanes', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['e2e'].get('rank0_ms_per_step'), 'roofline frac', j.get('roofline',{}).get('frac'))"
done
python - <<'PY' 2>&1 | tee $OUT/${TAG}_replay.txt
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__
cpbrt = __graft_entry__.build()
scene = cpbrt.Scene.load('scenes/cornel_box.pbrt'); scene.set_spp(1024)
r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
import torch
film = torch.zeros(900*900*3, dtype=torch.float32, device='cuda')
for rng in ((0, 1), (899, 900), (0, 1), (899, 900), (450, 451), (1023, 1024)):
    _, st = r.render(sample_begin=rng[0], sample_end=rng[1], device_film_ptr=film.data_ptr())
    print('samples', rng, 'ms_total %.3f' % st.ms_total)
PY
cut -c1-2500 $OUT/${TAG}_bench_lanes2.json
