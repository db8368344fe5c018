"""The reference's OWN renderer tests (tests/test_moderngl_renderer/test_volume_renderer.py: call plumbing into the
resource manager, TypeError messages, return types) run against the CUDA backend's VolumeRenderer + CudaManager,
and the INTEGRATION.md section 2 stub run verbatim against the reference's own Volume / Camera / Light /
RenderConfig classes.  Both need /root/reference (build container, no GPU), so the C ABI is a generated stub
library whose entry points all succeed (tools/run_reference_renderer_tests.py)."""

import json
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/tests"),
                                     reason="reference tree only exists in the build container")


@needs_reference
def test_reference_renderer_tests_pass_against_the_cuda_renderer():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_renderer_tests.py")],
                         capture_output=True, text=True, timeout=600)
    summary = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and summary["failed"] == [], (summary, out.stdout[-2000:])
    assert summary["passed"] >= 39, summary


def integration_stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# pyvr/cuda_renderer/__init__.py.*?)```", text, re.S)
    assert m, "INTEGRATION.md section 2 stub not found"
    return m.group(1)


DRIVER = r'''
import importlib.abc, importlib.machinery, sys, types
from unittest.mock import MagicMock

class _Stub(importlib.abc.MetaPathFinder, importlib.abc.Loader):       # moderngl / vtk / matplotlib are not installed
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in ("moderngl", "vtk", "matplotlib"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
    def create_module(self, spec):
        m = MagicMock(); m.__path__ = []; m.__name__ = spec.name; m.__spec__ = spec; return m
    def exec_module(self, module): pass
sys.meta_path.insert(0, _Stub())
sys.path.insert(0, "/root/reference")
import numpy as np
import pyvr
from pyvr.volume import Volume
from pyvr.camera import Camera
from pyvr.lighting import Light
from pyvr.config import RenderConfig
from pyvr.transferfunctions import ColorTransferFunction, OpacityTransferFunction
from pyvr.datasets import create_sample_volume

mod = types.ModuleType("pyvr.cuda_renderer"); mod.__package__ = "pyvr.cuda_renderer"; mod.__path__ = []
sys.modules["pyvr.cuda_renderer"] = mod
exec(compile(open(sys.argv[1]).read(), "INTEGRATION.md:stub", "exec"), mod.__dict__)
from pyvr.cuda_renderer import VolumeRenderer

r = VolumeRenderer(96, 64, config=RenderConfig.fast(), light=Light.directional([1, -1, 0]))
r.set_camera(Camera.isometric_view(distance=3.0))
r.load_volume(Volume(data=create_sample_volume(16, "sphere")))
ctf = ColorTransferFunction(control_points=[(0.0, (0, 0, 0)), (1.0, (1, 1, 1))])
r.set_transfer_functions(ctf, OpacityTransferFunction.linear(0.0, 0.3))
r.set_config(RenderConfig.balanced()); r.set_light(Light.default())
data = r.render()
assert isinstance(data, bytes) and len(data) == 96 * 64 * 4
for bad, call in ((1, r.load_volume), ("x", r.set_camera)):
    try:
        call(bad); raise SystemExit("no TypeError")
    except TypeError as e:
        assert str(e).startswith("Expected "), e
print("STUB-OK")
'''


@needs_reference
def test_integration_stub_runs_verbatim_against_the_reference_classes():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_reference_renderer_tests as runner

    with tempfile.TemporaryDirectory() as tmp:
        stub_lib = os.path.join(tmp, "libpyvr_cuda_stub.so")
        runner.stub_library(stub_lib)
        src = os.path.join(tmp, "stub.py")
        open(src, "w").write(integration_stub_source())
        drv = os.path.join(tmp, "driver.py")
        open(drv, "w").write(DRIVER)
        out = subprocess.run([sys.executable, drv, src], capture_output=True, text=True, timeout=300, cwd=tmp,
                             env={**os.environ, "PYVR_CUDA_LIB": stub_lib, "PYTHONDONTWRITEBYTECODE": "1"})
    assert out.returncode == 0 and "STUB-OK" in out.stdout, out.stderr[-3000:]


@pytest.mark.gpu
def test_integration_stub_renders_the_same_bytes_as_the_shipped_class():
    """The same stub, bound to the real library on the GPU box (the repo's mirror classes stand in for the
    reference's, which are not on that box): its frame equals pyvr_b200.cuda_renderer.VolumeRenderer's bit for bit."""
    import types

    import numpy as np

    import pyvr_b200
    from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig, Volume,
                           compute_normal_volume, create_sample_volume)
    from pyvr_b200.cuda_renderer import VolumeRenderer, _cabi

    os.environ.setdefault("PYVR_CUDA_LIB", _cabi.LIB_PATH)
    mod = types.ModuleType("pyvr_b200.integration_stub")
    mod.__package__ = "pyvr_b200.integration_stub"
    mod.__path__ = []
    sys.modules["pyvr_b200.integration_stub"] = mod
    exec(compile(integration_stub_source(), "INTEGRATION.md:stub", "exec"), mod.__dict__)
    data = create_sample_volume(64, "double_sphere")
    vol = Volume(data=data, normals=compute_normal_volume(data))
    cam, light, cfg = Camera.isometric_view(distance=3.0), Light.directional([1, -1, 0]), RenderConfig.balanced()
    ctf, otf = ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.3)
    frames = []
    for cls in (mod.VolumeRenderer, VolumeRenderer):
        r = cls(200, 160, config=cfg, light=light)
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_transfer_functions(ctf, otf)
        frames.append(np.frombuffer(r.render(), np.uint8))
    assert frames[0].any() and np.array_equal(frames[0], frames[1])
