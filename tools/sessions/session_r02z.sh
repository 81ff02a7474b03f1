#!/bin/bash
# round 2, 2-GPU session z: the 2-GPU parity test (four phases) and the per-rank time breakdown
OUT=gpurun_out
TAG=r02_z
python -m pytest tests/test_gpu_parity.py -x -q -k "two_gpu" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -12 $OUT/${TAG}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_breakdown.py 1024 > $OUT/${TAG}_breakdown.jsonl 2> $OUT/${TAG}_breakdown.err; echo "rc=$?"; cat $OUT/${TAG}_breakdown.jsonl; tail -3 $OUT/${TAG}_breakdown.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/multi_gpu_breakdown.py 256 >> $OUT/${TAG}_breakdown.jsonl 2>> $OUT/${TAG}_breakdown.err; tail -2 $OUT/${TAG}_breakdown.jsonl
