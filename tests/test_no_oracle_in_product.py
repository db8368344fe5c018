"""The product never routes through the oracle: nothing under pyvr_b200/ mentions it, and the
package has no numpy/CPU implementation of the march or of the normal stencil."""

import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_does_not_import_oracle():
    offenders = []
    for base, _, files in os.walk(os.path.join(ROOT, "pyvr_b200")):
        for name in files:
            if not name.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(base, name), errors="replace").read()
            if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle_render|oracle_normals", text, re.M):
                offenders.append(os.path.join(base, name))
    assert not offenders, offenders


def test_compute_entry_points_fail_loudly_without_gpu():
    import numpy as np
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the loud-failure path is for CPU-only boxes")
    from pyvr_b200 import compute_normal_volume
    from pyvr_b200.cuda_renderer import VolumeRenderer

    with pytest.raises(RuntimeError, match="pyvr_cuda error"):
        compute_normal_volume(np.zeros((4, 4, 4), np.float32))
    with pytest.raises(RuntimeError, match="pyvr_cuda error"):
        VolumeRenderer(32, 32)
