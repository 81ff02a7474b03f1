#!/bin/bash
# round 2, 8-GPU session v: bench.py at N=8 (+ reference arm on the same box's host cores), N=4, C4 and C5 on 8 GPUs
OUT=gpurun_out
TAG=r02_v
nvidia-smi -L | wc -l; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/${TAG}_n8.json 2> $OUT/${TAG}_n8.err; echo "N=8 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n4.json 2> $OUT/${TAG}_n4.err; echo "N=4 rc=$?"
python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 tools/config_multi_gpu.py --config C5 > $OUT/${TAG}_c5_n8.jsonl 2> $OUT/${TAG}_c5_n8.err; echo "C5 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tools/config_multi_gpu.py --config C4 > $OUT/${TAG}_c4_n8.jsonl 2> $OUT/${TAG}_c4_n8.err; echo "C4 rc=$?"
python - <<PY
import json
for n in (1,4,8):
    try:
        d=json.loads(open("$OUT/${TAG}_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(n, "failed", e)
try:
    d=json.loads(open("$OUT/${TAG}_reference.json").read().strip().splitlines()[-1]); print("reference", d["value"], d["cpu_baseline"]["cores"])
except Exception as e: print("reference failed", e)
PY
tail -c 1500 $OUT/${TAG}_c5_n8.jsonl; tail -c 1200 $OUT/${TAG}_c4_n8.jsonl
