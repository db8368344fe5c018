#!/usr/bin/env python
"""Run the REFERENCE's own renderer tests (tests/test_moderngl_renderer/test_volume_renderer.py) against the CUDA
backend's ``VolumeRenderer`` + ``CudaManager``.

    python tools/run_reference_renderer_tests.py [/root/reference]          (build container only)

Those tests never touch OpenGL: they patch ``moderngl.create_context`` and replace ``renderer.gl_manager.*`` with
mocks, then assert the call plumbing (which manager method is called with what), the ``TypeError`` messages and the
returned types.  Here a throw-away ``pyvr`` shim package maps ``pyvr.moderngl_renderer`` to
``pyvr_b200.cuda_renderer`` (nothing is copied from the reference), and -- the build container has no GPU -- the C
ABI is a stub library generated from ``_cabi.SYMBOLS`` whose entry points all succeed (``PYVR_CUDA_LIB``).  The
product itself never loads a stub: without ``libpyvr_cuda.so`` or a device it fails loudly.
Prints a one-line JSON summary and exits 0 iff every test passes.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REF = next((a for a in sys.argv[1:] if not a.startswith("-")), "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

SHIM = {
    "config.py": "config", "camera/__init__.py": "camera", "camera/camera.py": "camera", "camera/control.py": "camera",
    "lighting/__init__.py": "lighting", "lighting/light.py": "lighting", "volume/__init__.py": "volume",
    "volume/data.py": "volume", "transferfunctions/__init__.py": "transferfunctions",
    "transferfunctions/color.py": "transferfunctions", "transferfunctions/opacity.py": "transferfunctions",
    "transferfunctions/base.py": "transferfunctions", "datasets/__init__.py": "datasets",
}


def stub_library(path):
    from pyvr_b200.cuda_renderer import _cabi

    special = {
        "pyvr_cuda_create": "int pyvr_cuda_create(int d, int w, int h, void **out) { *out = (void *)1; return 0; }",
        "pyvr_cuda_abi_version": f"int pyvr_cuda_abi_version(void) {{ return {_cabi.ABI_VERSION}; }}",
        "pyvr_cuda_last_error": 'const char *pyvr_cuda_last_error(void) { return ""; }',
        "pyvr_cuda_host_alloc": "int pyvr_cuda_host_alloc(unsigned long n, void **out) { *out = calloc(n ? n : 1, 1); return *out ? 0 : -4; }",
        "pyvr_cuda_host_free": "int pyvr_cuda_host_free(void *p) { free(p); return 0; }",
    }
    src = ["#include <stdlib.h>"] + [special.get(n, f"int {n}() {{ return 0; }}") for n in _cabi.SYMBOLS]
    c = path[:-3] + ".c"
    open(c, "w").write("\n".join(src) + "\n")
    subprocess.run(["gcc", "-shared", "-fPIC", "-w", "-o", path, c], check=True)


def main():
    test_file = os.path.join(REF, "tests", "test_moderngl_renderer", "test_volume_renderer.py")
    if not os.path.exists(test_file):
        print(json.dumps({"skipped": f"{test_file} not found"}))
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        pkg = os.path.join(tmp, "shim", "pyvr")
        for rel, mod in SHIM.items():
            path = os.path.join(pkg, rel)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(f"import pyvr_b200.{mod} as _m\n"
                        "globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('__')})\n")
        open(os.path.join(pkg, "__init__.py"), "w").close()
        os.makedirs(os.path.join(pkg, "moderngl_renderer"))
        with open(os.path.join(pkg, "moderngl_renderer", "__init__.py"), "w") as f:
            f.write("from pyvr_b200.cuda_renderer import VolumeRenderer\nModernGLVolumeRenderer = VolumeRenderer\n")
        with open(os.path.join(pkg, "moderngl_renderer", "renderer.py"), "w") as f:
            f.write("from pyvr_b200.cuda_renderer.renderer import *\nfrom pyvr_b200.cuda_renderer import VolumeRenderer\n"
                    "ModernGLVolumeRenderer = VolumeRenderer\n")
        with open(os.path.join(pkg, "moderngl_renderer", "manager.py"), "w") as f:
            f.write("from pyvr_b200.cuda_renderer.manager import CudaManager as ModernGLManager\n")
        os.makedirs(os.path.join(tmp, "shim", "moderngl"))
        with open(os.path.join(tmp, "shim", "moderngl", "__init__.py"), "w") as f:   # the fixtures patch this name
            f.write("def create_context(*a, **k):\n    raise RuntimeError('no OpenGL in the CUDA backend')\n")
        stub = os.path.join(tmp, "libpyvr_cuda_stub.so")
        stub_library(stub)
        root = os.path.join(tmp, "root")
        os.makedirs(root)
        cmd = [sys.executable, "-m", "pytest", test_file, "-p", "no:cacheprovider", f"--rootdir={root}", "-q", "-rf"]
        env = {**os.environ, "PYTHONDONTWRITEBYTECODE": "1", "PYVR_CUDA_LIB": stub,
               "PYTHONPATH": os.path.join(tmp, "shim") + os.pathsep + REPO}
        out = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp, env=env)
    text = out.stdout + out.stderr
    failed = sorted(set(re.findall(r"FAILED \S+::(?:\w+::)?(\w+)", text)))
    m = re.search(r"(?:(\d+) failed, )?(\d+) passed", text)
    summary = {"passed": int(m.group(2)) if m else 0, "failed": failed}
    if "-v" in sys.argv:
        print(text[-6000:])
    print(json.dumps(summary))
    return 0 if m and not failed and "error" not in text.lower().split("passed")[-1] else 1


if __name__ == "__main__":
    sys.exit(main())
