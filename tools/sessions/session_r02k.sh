#!/bin/bash
# lanes x grid sweep (runtime knobs)
OUT=gpurun_out
run() { env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1))" | tee -a $OUT/r02_k_sweep.txt; }
run CPBRT_LANES=2
run CPBRT_LANES=3
run CPBRT_LANES=2 CPBRT_GRID_MATTE=3 CPBRT_GRID_TRACE=4
run CPBRT_LANES=2 CPBRT_GRID_MATTE=3
run CPBRT_LANES=2 CPBRT_GRID_MATTE=4 CPBRT_GRID_TRACE=6
run CPBRT_LANES=3 CPBRT_GRID_MATTE=3 CPBRT_GRID_TRACE=4
run CPBRT_LANES=2 CPBRT_GRID_TRACE=4
