#!/bin/bash
# round 2, 8-GPU session: the final build at N=8 and N=1 on the same box + the reference arm on its host cores (lean: one call)
OUT=gpurun_out
TAG=r02_final8b
nvidia-smi -L | wc -l; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n8.json 2> $OUT/${TAG}_n8.err; echo "N=8 rc=$?"
python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err
python - <<PY
import json
for n in (1,8):
    try:
        d=json.loads(open("$OUT/${TAG}_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(n, "failed", e)
try:
    d=json.loads(open("$OUT/${TAG}_reference.json").read().strip().splitlines()[-1]); print("reference", d["value"], d["cpu_baseline"]["cores"])
except Exception as e: print("reference failed", e)
PY
