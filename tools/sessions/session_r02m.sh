#!/bin/bash
OUT=gpurun_out
for lib in "" curry-pbrt_b200/build/variants/libcpbrt_base.so; do
  echo "== lib: ${lib:-current}"
  CPBRT_LIB=$lib python tools/config_bench.py --configs C4 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['msamples_s'],1), {k:round(v,1) for k,v in j['stage_ms'].items()}, 'build', round(j['bvh_build_s'],2))"
done
