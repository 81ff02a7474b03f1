#!/bin/bash
# round 2, GPU session y: a 128-spp share (one of eight ranks) with different pass sizes / lane counts, final build
OUT=gpurun_out
TAG=r02_y
run() { spp=$1; sp=$2; shift 2; env "$@" python bench.py --spp $spp --samples-per-pass $sp --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp pass $sp $*', round(j['value'],1), 'ms', round(j['ms_per_step'],2))" | tee -a $OUT/${TAG}_sweep.txt; }
run 1024 0 A=1
run 128 0 A=1
run 128 16 A=1
run 128 22 A=1
run 128 11 A=1
run 128 8 A=1
run 128 16 CPBRT_LANES=3
run 128 22 CPBRT_LANES=3
run 128 32 CPBRT_LANES=3
run 256 0 A=1
run 256 22 A=1
