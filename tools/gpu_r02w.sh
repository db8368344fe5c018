#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
c5() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 600 python bench.py --workload c5 --steps 8 --warmup 2 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C5(1 GPU, 2048^3, high_quality) $tag: %.3f ms/frame  march %.3f ms  %.1f Gsamples/s  fetched/s %.1f G' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['value'], d['roofline']['achieved']/64))" )
}
{
c5 "rows+pairs, 1 sample" PYVR_CUDA_BRICK8=0 PYVR_CUDA_TWO_SAMPLES=0 --
c5 "rows+pairs, 4 in flight" PYVR_CUDA_BRICK8=0 PYVR_CUDA_TWO_SAMPLES=1 --
c5 "bricks, 1 sample" PYVR_CUDA_BRICK8=1 PYVR_CUDA_TWO_SAMPLES=0 --
c5 "bricks, 4 in flight (auto)" X=0 --
} 2>&1 | tee $OUT/r02w_c5_ab.txt
