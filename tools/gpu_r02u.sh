#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
c4() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C4 $tag: %.2f ms/frame  %.1f Gsamples/s' % (d['ms_per_step'], d['value']))" )
}
L=$PWD/pyvr_b200/libpyvr_cuda
{
for v in k4b3 k4b4 k5b3 k6b3 k6b2 k8b2; do c4 "brick8 $v" PYVR_CUDA_BRICK8=1 PYVR_CUDA_LIB=${L}_$v.so --; done
for v in k4b3 k6b2; do echo "--- shard probe brick8 $v"; PYVR_CUDA_LIB=${L}_$v.so PYVR_CUDA_BRICK8=1 timeout 600 python tools/c4_shard_probe.py 2>&1 | grep -v Warning | grep -v "shift 0\|shift 2\|shift 4"; done
} 2>&1 | tee $OUT/r02u_ab.txt
