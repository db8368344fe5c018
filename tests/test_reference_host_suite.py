"""The reference's OWN unit tests for its host data model (config, camera, lighting, transfer functions), run
against this repo's mirror classes through a throw-away `pyvr` shim (tools/run_reference_host_tests.py).
Only possible where /root/reference exists (the build container); skipped on the GPU box."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference tree only exists in the build container")
def test_reference_unit_tests_pass_against_the_mirror_classes():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_host_tests.py")],
                         capture_output=True, text=True, timeout=600)
    summary = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and summary["unexpected_failures"] == [], summary
    assert summary["passed"] >= 141, summary
