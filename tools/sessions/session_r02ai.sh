#!/bin/bash
# round 2, GPU session ai: big-scene host path after its parallelisation -- full-size parity tests of C4 / C5, then their load times
OUT=gpurun_out
TAG=r02_ai
python -m pytest tests/test_gpu_config_sizes.py -x -q -k "c4 or c5" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
python tools/config_bench.py --configs C4,C5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['msamples_s'],1), 'parse', round(d['parse_s'],2), 'upload', round(d['upload_s'],3), 'build', round(d['bvh_build_s'],3), d['checks'])" | tee $OUT/${TAG}_configs.txt
