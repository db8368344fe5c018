import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    try:
        from pyvr_b200.cuda_renderer import _cabi

        return _cabi.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """GPU tests need the built library and a CUDA device: skip them (instead of failing inside
    pyvr_cuda_create) on a box that has neither.  `-m "not gpu"` is the CPU suite, `-m gpu` the parity suite."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device / libpyvr_cuda.so: the CUDA path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
