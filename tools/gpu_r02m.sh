#!/bin/bash
# round 2, call m: 2x2x2-texel brick layout (option "brick8"): parity with the row layouts, C4 and C3 A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q --tb=short -k "brick8 or layouts or half_texels or lut_sizes" 2>&1 | tail -8 | tee $OUT/r02m_pytest.txt
PYVR_CUDA_BRICK8=1 timeout 900 python -m pytest tests/test_render_gpu.py tests/test_sort_last_gpu.py tests/test_hwtex_gpu.py tests/test_synth_gpu.py tests/test_cuda_vs_shader_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -8 | tee -a $OUT/r02m_pytest.txt
timeout 600 python -m pytest tests/test_normals_gpu.py tests/test_synth_gpu.py -m gpu -q --tb=short 2>&1 | tail -8 | tee -a $OUT/r02m_pytest.txt
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
{
c4 "rows + z-pairs (brick8 0)" PYVR_CUDA_BRICK8=0 --
c4 "2x2x2 bricks (brick8 1)" PYVR_CUDA_BRICK8=1 --
c4 "rows unpaired" PYVR_CUDA_BRICK8=0 PYVR_CUDA_PAIR=0 --
c4 "auto" X=0 --
one "C3 f32 default" X=0 --
one "C3 f32 brick8" PYVR_CUDA_BRICK8=1 --
one "C3 f16 default" X=0 -- --texels f16
one "C3 f16 brick8" PYVR_CUDA_BRICK8=1 -- --texels f16
one "C3 f32 dense default" X=0 -- --no-ess
one "C3 f32 dense brick8" PYVR_CUDA_BRICK8=1 -- --no-ess
one "normals no-TMA" PYVR_NORMALS_NO_TMA=1 --
} 2>&1 | tee $OUT/r02m_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/r02m_c4_brick8 \
    env PYVR_CUDA_BRICK8=1 python bench.py --workload c4 --steps 1 --warmup 1 > $OUT/r02m_c4_brick8_ncu.log 2>&1
tail -1 $OUT/r02m_c4_brick8_ncu.log | cut -c1-150
