#!/bin/bash
# round 2, GPU session ab: observed image differences (device vs oracle) of every image-comparing test
OUT=gpurun_out
rm -f $OUT/r02_observed.txt
CPBRT_OBSERVE=$PWD/$OUT/r02_observed.txt python -m pytest tests/test_gpu_parity.py -x -q -k "not resident and not top_phase and not two_gpu and not checked" > $OUT/r02_ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_ab_pytest.log
sort $OUT/r02_observed.txt | awk '{ if (!($1 in m) || $2 > m[$1]) { m[$1] = $2; l[$1] = $3 } } END { for (k in m) print k, m[k], l[k] }' | sort
