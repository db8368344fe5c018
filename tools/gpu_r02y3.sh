#!/bin/bash
# round 2, call y3: unpaired rows x residues (8 slots per line), 4^3 cells, persistent tile queue -- combined, whole turntable
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
for swz in 5,2 3,6 6,3 5,6; do one "no z-pairs swz $swz" PYVR_CUDA_PAIR=0 PYVR_CUDA_SWZ=$swz --; done
one "no z-pairs + 4^3 cells" PYVR_CUDA_PAIR=0 PYVR_CUDA_LIB=${L}_cell4.so --
one "no z-pairs + persistent" PYVR_CUDA_PAIR=0 PYVR_CUDA_LIB=${L}_persist.so --
one "no z-pairs + 4^3 cells + persistent" PYVR_CUDA_PAIR=0 PYVR_CUDA_LIB=${L}_c4p.so --
one "z-pairs + 4^3 cells + persistent" PYVR_CUDA_LIB=${L}_c4p.so --
} 2>&1 | tee $OUT/r02y3_turntable_ab.txt
