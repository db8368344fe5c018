#!/usr/bin/env python
"""Run the REFERENCE's own unit tests for its host data model against this repo's mirror classes.

    python tools/run_reference_host_tests.py [/root/reference]          (build container only)

The reference's tests import `pyvr.camera`, `pyvr.config`, `pyvr.lighting`, `pyvr.transferfunctions`, ...; a
throw-away shim package named `pyvr` (written to a temp dir, nothing is copied from the reference) re-exports
the `pyvr_b200` mirrors under those names, and pytest is pointed at the reference's test files where they lie.
Out of scope and therefore excluded or expected to fail: the camera controllers / paths (`test_control.py`,
`test_trackball.py`, one light-linking integration test) and everything that needs matplotlib's colormap
registry (absent from this image; the reference fails the same five tests here).
Prints a one-line JSON summary and exits 0 iff only the expected tests fail.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPECTED_FAILURES = {
    "test_light_follows_camera_orbit",                      # needs CameraController (out of scope)
    "test_colormap_integration", "test_colormap_integration_success", "test_colormap_integration_import_error",
    "test_colormap_integration_invalid_name", "test_colormap_integration_other_errors",   # need matplotlib
}
SHIM = {
    "config.py": "config", "camera/__init__.py": "camera", "camera/camera.py": "camera", "camera/control.py": "camera",
    "lighting/__init__.py": "lighting", "lighting/light.py": "lighting", "volume/__init__.py": "volume",
    "volume/data.py": "volume", "transferfunctions/__init__.py": "transferfunctions",
    "transferfunctions/color.py": "transferfunctions", "transferfunctions/opacity.py": "transferfunctions",
    "transferfunctions/base.py": "transferfunctions", "datasets/__init__.py": "datasets",
}


def main():
    tests = os.path.join(REF, "tests")
    if not os.path.isdir(tests):
        print(json.dumps({"skipped": f"{tests} not found"}))
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
        root = os.path.join(tmp, "root")
        os.makedirs(root)
        targets = [os.path.join(tests, t) for t in ("test_config.py", "test_config_opacity_correction.py", "test_lighting",
                                                     "test_camera", "test_transferfunctions")]
        cmd = [sys.executable, "-m", "pytest", *targets, "-p", "no:cacheprovider", f"--rootdir={root}", "-q", "-rf",
               f"--ignore={tests}/test_camera/test_control.py", f"--ignore={tests}/test_camera/test_trackball.py"]
        env = {**os.environ, "PYTHONDONTWRITEBYTECODE": "1", "PYTHONPATH": os.path.join(tmp, "shim") + os.pathsep + REPO}
        out = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp, env=env)
    text = out.stdout + out.stderr
    failed = set(re.findall(r"FAILED \S+::(?:\w+::)?(\w+)", text))
    m = re.search(r"(?:(\d+) failed, )?(\d+) passed", text)
    summary = {"passed": int(m.group(2)) if m else 0, "failed": sorted(failed),
               "unexpected_failures": sorted(failed - EXPECTED_FAILURES)}
    print(json.dumps(summary))
    return 0 if m and not summary["unexpected_failures"] else 1


if __name__ == "__main__":
    sys.exit(main())
