#!/bin/bash
# round 2, GPU session aa: warp-redistributed exact phase of the flat search -- parity, then A/B against the per-lane loop
OUT=gpurun_out
TAG=r02_aa
python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sizes.py -x -q -k "not resident and not top_phase and not c5 and not million and not two_gpu" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
run() { spp=$1; shift; env "$@" python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
V=curry-pbrt_b200/build/variants
run 512 A=1
run 512 CPBRT_LIB=$V/libcpbrt_flat_per_lane_exact_phase.so
run 512 A=2
