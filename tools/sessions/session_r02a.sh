#!/bin/bash
# round 2, GPU session a: flat-search kernels -- parity tests, A/B bench against the loop schedule, ncu capture
set -x
OUT=gpurun_out
TAG=r02_a
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
CPBRT_SCHEDULE=loop python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_loop.json 2>> $OUT/${TAG}_bench.err
python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_trace_closest|k_trace_shadow|k_shade' --launch-skip 8 --launch-count 7 \
    -o $OUT/${TAG}_cornell python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_cornell.log 2>&1
cat $OUT/${TAG}_bench.json $OUT/${TAG}_bench_loop.json | cut -c1-1500
