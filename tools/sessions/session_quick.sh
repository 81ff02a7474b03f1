#!/bin/bash
# quick GPU check: parity tests + one bench line (+ optional extra commands passed as arguments)
set -x
OUT=gpurun_out
TAG=${1:-quick}
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-1200 $OUT/${TAG}_bench.json
