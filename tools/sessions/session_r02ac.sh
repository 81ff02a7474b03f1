#!/bin/bash
# round 2, GPU session ac: what the driver runs at round end -- the whole gpu suite, smoke(), bench.py with no flags, both arms
OUT=gpurun_out
TAG=r02_ac
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
( time python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2>&1 | grep real; cut -c1-400 $OUT/${TAG}_bench.json
( time python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err ) 2>&1 | grep real; cut -c1-300 $OUT/${TAG}_bench_reference.json
