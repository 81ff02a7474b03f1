#!/bin/bash
# round 2, GPU session e: restored Halton + pair tables: tests, bench, ncu capture of the Cornell kernels
set -x
OUT=gpurun_out
TAG=r02_e
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-900 $OUT/${TAG}_bench.json
python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_trace_closest|k_trace_shadow|k_shade|k_classify' --launch-skip 10 --launch-count 8 \
    -o $OUT/${TAG}_cornell python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_cornell.log 2>&1
