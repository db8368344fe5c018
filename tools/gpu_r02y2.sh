#!/bin/bash
# round 2, call y2: the remaining compile-time choices re-checked over the WHOLE turntable with the shipped lane arrangement
OUT=gpurun_out; mkdir -p $OUT
one() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 300 python bench.py --steps 2 --warmup 1 --views-per-step 180 --skip-cpu-baseline --no-alternatives "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag: value=%.1f Gs/s ms/view=%.4f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))" )
}
L=$PWD/pyvr_b200/libpyvr_cuda
{
one "shipped (2x2 passes, 3,2, 7 CTAs/SM, 8^3 cells)" X=0 --
for swz in 2,3 1,2 2,2 3,0; do one "swz $swz" PYVR_CUDA_SWZ=$swz --; done
for v in mb6 mb8 cell4 persist lutp; do one "$v" PYVR_CUDA_LIB=${L}_$v.so --; done
one "no z-pairs" PYVR_CUDA_PAIR=0 --
} 2>&1 | tee $OUT/r02y2_turntable_ab.txt
