#!/bin/bash
# A/B of library builds and layout options on the headline workload (device-resident value only).
# Usage: bash tools/gpu_ab.sh "<lib suffix list>" "<pair option list>" [extra bench flags]
LIBS=${1:-"default"}; PAIRS=${2:-"-1"}; shift; shift
for lib in $LIBS; do for pair in $PAIRS; do
  if [ "$lib" != "default" ]; then export PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_$lib.so; else unset PYVR_CUDA_LIB; fi
  export PYVR_CUDA_PAIR=$pair
  timeout 300 python bench.py --steps 3 --warmup 3 --views-per-step 6 --skip-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('lib=$lib pair=$pair $*: value=%.1f Gs/s fps=%.1f e2e=%.1f ms/launch=%.2f' % (d['value'], d['frames_per_s'], d['e2e']['value'], r['kernel_ms_per_launch']))"
done; done
