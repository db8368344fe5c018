#!/bin/bash
# round 2, call q: two samples in flight per ray (f16x4 z-pairs): bit-identity, C4 / C3-f16 A/B at 6..9 CTAs per SM, shard probe
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q --tb=short -k "two_samples or half_texels" 2>&1 | tail -6 | tee $OUT/r02q_pytest.txt
PYVR_CUDA_TWO_SAMPLES=1 timeout 900 python -m pytest tests/test_render_gpu.py tests/test_sort_last_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -4 | tee -a $OUT/r02q_pytest.txt
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
L=$PWD/pyvr_b200/libpyvr_cuda
{
c4 "one sample per iteration" PYVR_CUDA_TWO_SAMPLES=0 --
c4 "two samples, 7 CTAs/SM (auto)" X=0 --
for v in two6 two8 two9; do c4 "two samples $v" PYVR_CUDA_LIB=${L}_$v.so --; done
one "C3 f16 one sample" PYVR_CUDA_TWO_SAMPLES=0 -- --texels f16
one "C3 f16 two samples 7" PYVR_CUDA_TWO_SAMPLES=1 -- --texels f16
one "C3 f16 two samples 9" PYVR_CUDA_TWO_SAMPLES=1 PYVR_CUDA_LIB=${L}_two9.so -- --texels f16
} 2>&1 | tee $OUT/r02q_ab.txt
timeout 900 python tools/c4_shard_probe.py 2>&1 | grep -v Warning | tee $OUT/r02q_c4_shard_probe.txt
