#!/bin/bash
# ncu --set full capture of the march kernel for the given bench flags.  Usage: bash tools/gpu_ncu.sh TAG [bench flags...]
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG} \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline "$@" > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log | cut -c1-300
