#!/bin/bash
# round 2, GPU session n (three calls: outputs of one call must stay below 64 MiB)
#   tools/session_r02n.sh bench    bench line + reference arm, C4 traversal roofline, configs C1..C4, ncu launch list
#   tools/session_r02n.sh cornell  ncu --set full capture of the Cornell kernels (one 32-spp pass, second bounce)
#   tools/session_r02n.sh c4       ncu --set full capture of the C4 traversal kernels
set -x
OUT=gpurun_out
TAG=r02_n
case "$1" in
bench)
python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python tools/traversal_roofline.py --rays 22 > $OUT/${TAG}_traversal_c4.jsonl 2>> $OUT/${TAG}_bench.err
python tools/config_bench.py --configs C1,C2,C3,C4 --cpu > $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_bench.err
python bench.py --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
;;
cornell)
python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_trace_closest|k_trace_shadow|k_shade|k_classify' --launch-skip 14 --launch-count 7 \
    -o $OUT/${TAG}_cornell python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_cornell.log 2>&1
;;
c4)
python tools/traversal_roofline.py --rays 20 --no-oracle --repeat 1 > $OUT/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_batch_closest|k_batch_any' --launch-skip 9 --launch-count 7 \
    -o $OUT/${TAG}_c4 python tools/traversal_roofline.py --rays 20 --no-oracle --repeat 1 > $OUT/${TAG}_ncu_c4.log 2>&1
;;
esac
ls -la $OUT | tail -14
