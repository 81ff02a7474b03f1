#!/bin/bash
# round 2, GPU session r: pass-size sweep at 1024 spp and at 128 spp (= what one of eight ranks renders)
OUT=gpurun_out
TAG=r02_r
run() { spp=$1; sp=$2; shift 2; env "$@" python bench.py --spp $spp --samples-per-pass $sp --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp pass $sp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'ms', round(j['ms_per_step'],2))" | tee -a $OUT/${TAG}_sweep.txt; }
run 1024 32 A=1
run 1024 16 A=1
run 1024 24 A=1
run 1024 48 A=1
run 1024 64 A=1
run 128 32 A=1
run 128 16 A=1
run 128 8 A=1
run 128 16 CPBRT_LANES=3
run 128 32 CPBRT_LANES=1
