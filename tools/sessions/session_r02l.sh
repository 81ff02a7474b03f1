#!/bin/bash
# round 2, GPU session l: tight-only search with end-of-ray reachability proof: all parity tests, C4 numbers
set -x
OUT=gpurun_out
TAG=r02_l
python -m pytest tests -x -q -m gpu --durations=5 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
python tools/traversal_roofline.py --rays 22 --no-oracle > $OUT/${TAG}_traversal_c4.jsonl 2> $OUT/${TAG}_trav.err; cut -c1-400 $OUT/${TAG}_traversal_c4.jsonl
python tools/config_bench.py --configs C2,C4 > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-700 $OUT/${TAG}_configs.jsonl
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
j=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]); print(round(j['value'],1), 'e2e', round(j['e2e']['value'],1)); r=j['roofline']; print({k:r[k] for k in ('achieved','peak','frac','grays_s','own_node_steps_per_ray','own_prim_tests_per_ray','avg_launch_ms')}); print(j['roofline_shading_stage'])"
