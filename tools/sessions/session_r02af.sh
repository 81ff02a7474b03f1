#!/bin/bash
# round 2, GPU session af: edge-phase unroll factor, repeated (is the difference noise?)
OUT=gpurun_out
TAG=r02_af
run() { spp=$1; shift; env "$@" python bench.py --spp $spp --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
V=curry-pbrt_b200/build/variants
for rep in 1 2; do
run 1024 A=$rep
run 1024 CPBRT_LIB=$V/libcpbrt_flat_unroll_1.so
run 1024 CPBRT_LIB=$V/libcpbrt_flat_unroll_2.so
done
