#!/bin/bash
# round 2, GPU session x: new defaults (Cornell), vote-kernel block variants on C4, specular block variants
OUT=gpurun_out
TAG=r02_x
run() { spp=$1; shift; env "$@" python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
c4() { env "$@" python tools/config_bench.py --configs C4 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 $*', round(j['msamples_s'],1), {k:round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
V=curry-pbrt_b200/build/variants
run 512 A=1
run 512 CPBRT_LIB=$V/libcpbrt_sp7.so
run 512 CPBRT_LIB=$V/libcpbrt_sp6.so
c4 A=1
for v in vote_c5_s6 vote_c5_s8 vote_c6_s6 vote_c4_s6 vote_c7_s8; do c4 CPBRT_LIB=$V/libcpbrt_$v.so; done
c4 CPBRT_LANES=3
