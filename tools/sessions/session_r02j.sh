#!/bin/bash
OUT=gpurun_out
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/r02_j_bench.json 2> $OUT/r02_j_bench.err; echo "rc=$?"
python -c "
import json
j=json.loads(open('$OUT/r02_j_bench.json').read().strip().splitlines()[-1]); print(round(j['value'],1), j['rank0_device_ms_of_each_step'], j['clocks']); e=j['e2e']; print('e2e', round(e['value'],1), e['rank0_ms_of_each_step'], e['rank0_device_ms_of_each_step'], e['clocks'], e['h2d_bytes_per_step'])"
