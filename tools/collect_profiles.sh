#!/bin/bash
# One GPU-box session that produces every measured artefact kept under profiles/ (run via gpurun):
#   bench lines (own arm + reference arm), ncu launch list, ncu --set full captures of the top
#   kernels on the Cornell box and on the 1M-triangle traversal benchmark, traversal roofline.
# Numbers printed under ncu are never used as bench values.
set -x
OUT=gpurun_out
TAG=${1:-r01_final}
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python tools/traversal_roofline.py --rays 22 > $OUT/${TAG}_traversal_c4.jsonl 2>> $OUT/${TAG}_bench.err
python tools/config_bench.py --configs C1,C2,C3,C4 --cpu > $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_bench.err
python tools/bvh_build_bench.py --min-lens 256 > $OUT/${TAG}_bvh_build.jsonl 2>> $OUT/${TAG}_bench.err
python tools/schedule_crossover.py > $OUT/${TAG}_schedule_crossover.jsonl 2>> $OUT/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace_closest|k_trace_shadow|k_shade' --launch-skip 8 --launch-count 7 \
    -o $OUT/${TAG}_cornell python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_cornell.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_batch_closest|k_batch_any' --launch-skip 6 --launch-count 6 \
    -o $OUT/${TAG}_c4 python tools/traversal_roofline.py --rays 20 --no-oracle --repeat 1 > $OUT/${TAG}_ncu_c4.log 2>&1
ls -la $OUT | tail -12
