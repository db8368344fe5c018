#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q --tb=short -k "two_samples" 2>&1 | tail -3 | tee $OUT/r02s_pytest.txt
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
c4 "K=2, 5 CTAs/SM (default)" X=0 --
for v in k3b5 k3b4 k4b4 k4b3; do c4 "$v" PYVR_CUDA_LIB=${L}_$v.so --; done
for v in k3b4 k4b4; do echo "--- shard probe $v"; PYVR_CUDA_LIB=${L}_$v.so timeout 600 python tools/c4_shard_probe.py 2>&1 | grep -v Warning | grep -v "shift 0\|shift 2\|shift 4"; done
} 2>&1 | tee $OUT/r02s_ab.txt
