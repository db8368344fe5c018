#!/bin/bash
# round 2, call n: K2 with the branch-free finish at 4 / 5 / 6 CTAs per SM (+ ncu); brick8 on C4 with L2 / L1 prefetch and more warps
OUT=gpurun_out; mkdir -p $OUT
c4() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C4 $tag: %.2f ms/frame  %.1f Gsamples/s' % (d['ms_per_step'], d['value']))" )
}
one() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag: value=%.1f Gs/s ms/view=%.3f normals_ms=%.4f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch'], d['normals_kernel'].get('kernel_ms', -1)))" )
}
L=$PWD/pyvr_b200/libpyvr_cuda
{
one "K2 5 CTAs/SM (default)" X=0 --
one "K2 4 CTAs/SM" PYVR_CUDA_LIB=${L}_nb4.so --
one "K2 6 CTAs/SM" PYVR_CUDA_LIB=${L}_nb6.so --
c4 "brick8" PYVR_CUDA_BRICK8=1 --
for v in b8pf2 b8pf4 b8pf8 b8pf4l1 b8occ11; do c4 "brick8 $v" PYVR_CUDA_BRICK8=1 PYVR_CUDA_LIB=${L}_$v.so --; done
c4 "rows + z-pairs" PYVR_CUDA_BRICK8=0 --
} 2>&1 | tee $OUT/r02n_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^normals_" -s 2 -c 1 -f -o $OUT/r02n_normals \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives > $OUT/r02n_normals_ncu.log 2>&1
tail -1 $OUT/r02n_normals_ncu.log | cut -c1-150
