#!/bin/bash
# round 2, call h: normals tests (shuffle fix), K2 timing, C4 on one GPU with lane-arrangement / pair / rotation variants
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_normals_gpu.py tests/test_render_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
python - <<'PY'
import ctypes, numpy as np, torch, sys
sys.path.insert(0, '.')
from pyvr_b200.cuda_renderer import _cabi
from pyvr_b200 import create_sample_volume
for n in (512, 768):
    d_in = torch.rand((n, n, n), device='cuda'); d_out = torch.empty((n, n, n, 3), device='cuda')
    ms, best = ctypes.c_float(0), 1e9
    for _ in range(8):
        _cabi.check(_cabi.lib().pyvr_cuda_compute_normals(0, ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_out.data_ptr()), n, n, n, 1, ctypes.byref(ms)))
        best = min(best, ms.value)
    print(f'K2 {n}^3: {best:.4f} ms  {n**3*16/best/1e6:.0f} GB/s  frac {n**3*16/best/1e6/6550.1:.3f}')
    del d_in, d_out
PY
c4() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C4 $tag: %.2f ms/frame  %.1f Gsamples/s  fetched share %.3f' % (d['ms_per_step'], d['value'], d['roofline']['achieved']*1e9/64/ (d['value']*1e9) ))" )
}
{
c4 "default (arr1, pair auto)" X=0 --
c4 "arr0" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_arr0.so --
c4 "arr0 swz 1,3" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_arr0.so PYVR_CUDA_SWZ=1,3 --
c4 "default pair0" PYVR_CUDA_PAIR=0 --
c4 "default linear rows" PYVR_CUDA_LAYOUT=linear --
c4 "hwtex" X=0 -- --hwtex
} 2>&1 | tee $OUT/r02h_c4_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 2 -c 1 -f -o $OUT/r02h_c4_march \
    python bench.py --workload c4 --steps 1 --warmup 2 > $OUT/r02h_c4_ncu.log 2>&1
tail -1 $OUT/r02h_c4_ncu.log | cut -c1-150
