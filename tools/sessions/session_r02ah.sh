#!/bin/bash
# round 2, GPU session ah: cold 64-bit Halton paths and the homogeneous divide out of line -- parity, then A/B
OUT=gpurun_out
TAG=r02_ah
python -m pytest tests/test_gpu_parity.py -x -q -k "halton or race_free or full_path or general or camera or thin_lens or first_hit or directlighting" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
run() { spp=$1; shift; env "$@" python bench.py --spp $spp --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spp $spp $*', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {k: round(v,1) for k,v in j['stage_ms'].items()})" | tee -a $OUT/${TAG}_sweep.txt; }
V=curry-pbrt_b200/build/variants
for rep in 1 2; do
run 1024 A=$rep
run 1024 CPBRT_LIB=$V/libcpbrt_inline_cold_paths.so
done
