#!/bin/bash
# round 2, GPU session ag: block counts around the unroll-1 edge phase
OUT=gpurun_out
TAG=r02_ag
run() { spp=$1; shift; env "$@" python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
V=curry-pbrt_b200/build/variants
for v in flat_unroll_1 u1_c6 u1_c4 u1_s5 u1_s7 u1_m5 u1_m7; do run 512 CPBRT_LIB=$V/libcpbrt_$v.so; done
run 512 CPBRT_LIB=$V/libcpbrt_flat_unroll_1.so CPBRT_LANES=3
