#!/bin/bash
# Halton chunk-table bound sweep (runtime knob): shade-stage time per setting
OUT=gpurun_out
python -m pytest tests -m gpu -x -q -k "halton or full_path" > $OUT/r02_d_pytest.log 2>&1; tail -2 $OUT/r02_d_pytest.log
for m in 0 64 256 1024 4096 65536; do
  CPBRT_HALTON_CHUNK_MAX=$m python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_max $m', round(j['value'],1), {k:round(v,1) for k,v in j['stage_ms'].items()}, 'e2e', round(j['e2e']['value'],1))" | tee -a $OUT/r02_d_halton_sweep.txt
done
