#!/bin/bash
# K2 timing A/B (CUDA events inside pyvr_cuda_compute_normals) + tests.
python -m pytest tests/test_normals_gpu.py -m gpu -q 2>&1 | tail -2
python - <<'PY'
import os, numpy as np, sys
sys.path.insert(0, '.')
from pyvr_b200 import create_sample_volume
from pyvr_b200.cuda_renderer import _cabi
import torch
for n in (256, 512, 768):
    v = torch.rand((n, n, n), device='cuda', dtype=torch.float32)
    out = torch.empty((n, n, n, 3), device='cuda', dtype=torch.float32)
    import ctypes
    ms = ctypes.c_float()
    best = 1e9
    for it in range(8):
        _cabi.check(_cabi.lib().pyvr_cuda_compute_normals(0, ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(out.data_ptr()), n, n, n, 1, ctypes.byref(ms)))
        best = min(best, ms.value)
    print(f"n={n} v1={os.environ.get('PYVR_NORMALS_V1')} best {best:.3f} ms  {n**3*16/best/1e6:.0f} GB/s algorithmic")
PY
