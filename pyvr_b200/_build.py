"""In-tree build of ``libpyvr_cuda.so`` with nvcc for sm_100a (no JIT cache, no setup.py).

    python -m pyvr_b200._build [--force]

Flags: ``-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false``.  ``-fmad=false`` is
part of the arithmetic contract: the compiler never contracts ``a*b+c`` itself, the
kernels spell every fused operation as ``fmaf``.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libpyvr_cuda.so")
SOURCES = ["abi.cu", "march.cu", "volume_pack.cu", "normals.cu", "composite.cu", "synth.cu", "bandwidth.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(REPO, "include", "pyvr_cuda.h")]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), suffix: str = "") -> str:
    """Compile the CUDA library if it is missing or older than its sources; returns its path.

    ``defines`` / ``suffix`` build an experiment variant ``libpyvr_cuda_<suffix>.so`` with extra ``-D`` macros
    (tools/gpu_ab.sh selects one with ``PYVR_CUDA_LIB``); the product is always the plain build."""
    out = LIB_PATH if not suffix else os.path.join(PKG_DIR, f"libpyvr_cuda_{suffix}.so")
    if not suffix and not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(REPO, "include"), "-I", CSRC, *[f"-D{d}" for d in defines]]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out


if __name__ == "__main__":
    # python -m pyvr_b200._build [--force] [-v] [--variant NAME -DMACRO=1 ...]
    argv = sys.argv[1:]
    variant = argv[argv.index("--variant") + 1] if "--variant" in argv else ""
    print(build_library(force="--force" in argv, verbose="-v" in argv,
                        defines=[a[2:] for a in argv if a.startswith("-D")], suffix=variant))
