"""The C-ABI library builds, loads, and exports every symbol include/pyvr_cuda.h declares (no GPU needed)."""

import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pyvr_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pyvr_cuda_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pyvr_b200 import _build
    from pyvr_b200.cuda_renderer import _cabi

    _build.build_library()
    declared = _declared()
    assert len(declared) >= 18
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (pyvr_cuda_\w+)", out.stdout))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(_cabi.SYMBOLS), (sorted(set(declared) ^ set(_cabi.SYMBOLS)))
    lib = _cabi.lib()
    assert lib.pyvr_cuda_abi_version() == _cabi.ABI_VERSION


def test_library_is_sm100a_only():
    from pyvr_b200.cuda_renderer import _cabi

    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_view_from_matrices_matches_closed_form():
    """pyvr_cuda_view_from_matrices (host-only code): u = right*aspect*tan(fov/2), v = up*tan(fov/2), w = forward."""
    from pyvr_b200 import Camera
    from pyvr_b200.cuda_renderer import _cabi

    for cam, aspect in ((Camera.isometric_view(distance=3.0), 1.0),
                        (Camera(azimuth=1.1, elevation=-0.4, roll=0.7, distance=2.2, fov=0.9), 16 / 9)):
        v = _cabi.view_from_camera(cam, aspect)
        pos, up = cam.get_camera_vectors()
        fwd = cam.target - pos
        fwd = fwd / np.linalg.norm(fwd)
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        true_up = np.cross(right, fwd)
        t = np.tan(cam.fov / 2)
        np.testing.assert_allclose(list(v.origin), pos, atol=1e-6)
        np.testing.assert_allclose(list(v.u), right * t * aspect, atol=2e-6)
        np.testing.assert_allclose(list(v.v), true_up * t, atol=2e-6)
        np.testing.assert_allclose(list(v.w), fwd, atol=2e-6)


def test_error_reporting_without_gpu():
    """Bad arguments are rejected by host-side validation with a message (no device work)."""
    import ctypes

    from pyvr_b200.cuda_renderer import _cabi

    lib = _cabi.lib()
    assert lib.pyvr_cuda_create(0, 16, 16, None) == -1
    assert "NULL" in _cabi.last_error()
    out = _cabi.View()
    z = (ctypes.c_float * 16)()
    pos = (ctypes.c_float * 3)()
    assert lib.pyvr_cuda_view_from_matrices(z, z, pos, ctypes.byref(out)) == -1
    assert "singular" in _cabi.last_error()


def test_renderer_type_errors_before_device_work():
    """Constructor argument checks of the reference (renderer.py:77-78, 85-86) fire before any CUDA call."""
    import pytest

    from pyvr_b200.cuda_renderer import VolumeRenderer

    with pytest.raises(TypeError, match="Expected RenderConfig instance"):
        VolumeRenderer(config="fast")
    with pytest.raises(TypeError, match="Expected Light instance"):
        VolumeRenderer(light=object())
