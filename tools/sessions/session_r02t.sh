#!/bin/bash
# round 2, GPU session t: fused shade-and-trace kernels -- parity, then flat vs fused and the block-count variants
OUT=gpurun_out
TAG=r02_t
python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sizes.py -x -q -k "not resident and not top_phase and not c5 and not million" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
run() { env "$@" python bench.py --spp 256 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
run CPBRT_SCHEDULE=flat
run CPBRT_SCHEDULE=fused
V=curry-pbrt_b200/build/variants
run CPBRT_SCHEDULE=fused CPBRT_LIB=$V/libcpbrt_fused_matte_4.so
run CPBRT_SCHEDULE=fused CPBRT_LIB=$V/libcpbrt_fused_matte_3.so
run CPBRT_SCHEDULE=fused CPBRT_LIB=$V/libcpbrt_fused_matte_6.so
run CPBRT_SCHEDULE=fused CPBRT_LANES=1
run CPBRT_SCHEDULE=fused CPBRT_LANES=3
