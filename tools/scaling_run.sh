#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on one box, the way the driver launches it; one JSON line per N.
OUT=gpurun_out
TAG=${1:-scale}
nvidia-smi -L | wc -l
python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 2 --warmup 3 > $OUT/${TAG}_n$N.json 2> $OUT/${TAG}_n$N.err
  echo "N=$N rc=$?"
done
python - <<PY
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open("$OUT/${TAG}_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(n, "failed", e)
PY
