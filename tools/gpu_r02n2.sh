#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_normals_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6 | tee $OUT/r02n2_pytest.txt
timeout 120 python - <<'PY' 2>&1 | tee $OUT/r02n2_e2e.txt
import time, os, numpy as np
from pyvr_b200 import create_sample_volume
from pyvr_b200.cuda_renderer import _cabi
import oracle
d = create_sample_volume(512, "double_sphere")
for tag, env in (("pipeline", None), ):
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter(); n = _cabi.compute_normals_host(d); best = min(best, time.perf_counter() - t0)
    print(tag, "512^3 host->host best of 4: %.3f s" % best)
want = oracle.normals(d)
print("bit-identical to the oracle:", bool(np.array_equal(n.view(np.uint32), want.view(np.uint32))))
PY
