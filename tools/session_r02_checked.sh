#!/bin/bash
# the bounds-checked build of the library over the every-kernel workload (compute-sanitizer is closed on this pool)
OUT=gpurun_out
CPBRT_LIB=curry-pbrt_b200/build/variants/libcpbrt_checked.so python tools/sanitize_workload.py > $OUT/r02_checked_workload.txt 2>&1; echo "rc=$?" >> $OUT/r02_checked_workload.txt
tail -32 $OUT/r02_checked_workload.txt
