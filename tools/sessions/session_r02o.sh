#!/bin/bash
OUT=gpurun_out
TAG=r02_o
CPBRT_TIMING=1 python tools/bvh_build_bench.py --sizes 2500x2000 --min-lens 256 > $OUT/${TAG}_bvh_build.jsonl 2> $OUT/${TAG}_bvh_build.err; cat $OUT/${TAG}_bvh_build.jsonl | cut -c1-600; grep "cpbrt_build_bvh\|gpu_sah" $OUT/${TAG}_bvh_build.err | tail -22
