#!/bin/bash
# round 2, GPU session w: block-count variants of the flat trace kernels and the matte kernel
OUT=gpurun_out
TAG=r02_w
run() { spp=$1; shift; env "$@" python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
V=curry-pbrt_b200/build/variants
for v in c5_s6_m6 c5_s6_m7 c5_s5_m7 c5_s4_m6 c5_s5_m6_sp5 c5_s7_m6; do run 512 CPBRT_LIB=$V/libcpbrt_$v.so; done
run 512 CPBRT_LIB=$V/libcpbrt_c5_s6_m6.so CPBRT_LANES=3
