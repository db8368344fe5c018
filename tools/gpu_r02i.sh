#!/bin/bash
# round 2, call i: C4 prefetch distances + round-1 build on the same box; C3 with interval-table sizes
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
print('$tag: value=%.1f Gs/s ms/view=%.3f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))" )
}
{
c4 "auto prefetch" X=0 --
for pf in 0 2 4 8 16; do c4 "prefetch $pf" PYVR_CUDA_PREFETCH=$pf --; done
( cd _r01 && timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C4 round-1 build (b367ccb): %.2f ms/frame  %.1f Gsamples/s' % (d['ms_per_step'], d['value']))" )
one "C3 default" X=0 --
one "C3 prefetch 4 (forced)" PYVR_CUDA_PREFETCH=4 --
one "C3 iv4" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_iv4.so --
one "C3 iv3" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_iv3.so --
( cd _r01 && timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('C3 round-1 build (b367ccb): value=%.1f Gs/s ms/view=%.3f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))" )
} 2>&1 | tee $OUT/r02i_ab.txt
