#!/bin/bash
# round 2, call k: lane arrangements whose 4-lane data-stage pass is a 2x2 pixel block (rule measured by
# tools/l1_gather_probe.cu) x pitch residues; persistent tile queue
OUT=gpurun_out; mkdir -p $OUT
one() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag: value=%.1f Gs/s ms/view=%.3f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))" )
}
{
one "arr1 (default) swz 3,1 ess" X=0 --
one "arr1 (default) swz 3,1 dense" X=0 -- --no-ess
for arr in arr3 arr4; do
  for swz in 3,1 2,2 2,1 2,3 1,1 3,2; do
    one "$arr swz $swz ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_$arr.so PYVR_CUDA_SWZ=$swz --
  done
  one "$arr swz 2,2 dense" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_$arr.so PYVR_CUDA_SWZ=2,2 -- --no-ess
  one "$arr swz 2,1 dense" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_$arr.so PYVR_CUDA_SWZ=2,1 -- --no-ess
done
one "arr1 swz 2,2 ess" PYVR_CUDA_SWZ=2,2 --
one "persist ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_persist.so --
one "persist dense" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_persist.so -- --no-ess
one "arr3 f16 swz 3,1 ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_arr3.so -- --texels f16
one "arr1 f16 swz 3,1 ess" X=0 -- --texels f16
one "arr1 f16 swz 6,2 ess" PYVR_CUDA_SWZ=6,2 -- --texels f16
one "arr4 f16 swz 4,3 ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_arr4.so PYVR_CUDA_SWZ=4,3 -- --texels f16
} 2>&1 | tee $OUT/r02k_ab.txt
