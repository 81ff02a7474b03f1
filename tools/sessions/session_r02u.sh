#!/bin/bash
# round 2, 2-GPU session u: the 2-GPU parity test, then bench.py at N=1 and N=2 the way the driver launches it
OUT=gpurun_out
TAG=r02_u
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_parity.py -x -q -k "two_gpu" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -12 $OUT/${TAG}_pytest.log
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/${TAG}_n2.json 2> $OUT/${TAG}_n2.err; echo "N=2 rc=$?"
python - <<PY
import json
for n in (1,2):
    try:
        d=json.loads(open("$OUT/${TAG}_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(n, "failed", e)
PY
