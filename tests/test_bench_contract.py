"""bench.py's command-line contract, checked without a GPU: the reference arm prints one well-formed JSON
line (it is the CPU path), and the default arm refuses to run without a CUDA device instead of falling back."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                          timeout=600, env={**os.environ, **(env or {})})


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--size", "48", "--width", "96", "--height", "64", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gsamples/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["value"] > 0
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    # "reference" = the reference's GLSL on Mesa llvmpipe (oracle/gl); "port" = the C restatement when Mesa is missing
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "view" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit that (round-1 SCALE runs did)."""
    out = _run("--impl", "reference", "--size", "32", "--width", "64", "--height", "48", "--steps", "1", "--warmup", "0",
               "--reference-backend", "oracle", env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == cores


def test_reference_arm_other_ranks_exit_quietly():
    out = _run("--impl", "reference", "--size", "32", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_default_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    out = _run("--steps", "1", "--warmup", "0", "--size", "32", "--width", "64", "--height", "48")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
