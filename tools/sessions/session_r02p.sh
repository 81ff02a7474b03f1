#!/bin/bash
# round 2, GPU session p: device-resident SAH build (tests + timing), schedule race test
OUT=gpurun_out
TAG=r02_p
python -m pytest tests/test_gpu_parity.py -x -q -k "resident or top_phase or bvh_order or race_free" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -15 $OUT/${TAG}_pytest.log
for i in 1 2 3; do CPBRT_TIMING=1 python tools/bvh_build_bench.py --sizes 2500x2000 --min-lens 256 > $OUT/${TAG}_bvh_build$i.jsonl 2> $OUT/${TAG}_bvh_build$i.err; cat $OUT/${TAG}_bvh_build$i.jsonl | cut -c1-900; done
grep "cpbrt_build_bvh\|gpu_sah" $OUT/${TAG}_bvh_build3.err | sed -n 5,20p
CPBRT_TIMING=1 python tools/bvh_build_bench.py --sizes 1000x500 --min-lens 256 > $OUT/${TAG}_bvh_build.jsonl 2> $OUT/${TAG}_bvh_build.err; cat $OUT/${TAG}_bvh_build.jsonl | cut -c1-900
cat $OUT/${TAG}_bvh_build1.jsonl >> $OUT/${TAG}_bvh_build.jsonl
