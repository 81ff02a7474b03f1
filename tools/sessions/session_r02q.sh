#!/bin/bash
# round 2, GPU session q: Halton digit-group tables -- parity, then the group-size sweep on the Cornell bench
OUT=gpurun_out
TAG=r02_q
python -m pytest tests/test_gpu_parity.py -x -q -k "halton or full_path or first_hit" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
for g in 2 128 1024 4096 16384; do
  CPBRT_HALTON_GROUP=$g python bench.py --spp 256 --steps 3 --warmup 3 --no-cpu-baseline 2> $OUT/${TAG}_bench_$g.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('group_max $g', round(j['value'],1), 'Msamples/s', {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt
done
