#!/bin/bash
# round 2, call y: pitch residues and lane arrangements over the WHOLE turntable (the earlier A/Bs marched views 0..15)
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
for rx in 0 1 2 3; do for ry in 0 1 2 3; do one "arr1 swz $rx,$ry" PYVR_CUDA_SWZ=$rx,$ry --; done; done
for arr in arr0 arr3 arr4; do for swz in 3,1 1,3 2,1 3,2; do one "$arr swz $swz" PYVR_CUDA_LIB=${L}_$arr.so PYVR_CUDA_SWZ=$swz --; done; done
} 2>&1 | tee $OUT/r02y_turntable_ab.txt
