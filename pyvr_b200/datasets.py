"""Synthetic test volumes (host) and the normal-volume stencil (device).

``create_sample_volume`` restates the analytic shapes of the reference's input
generator (``pyvr/datasets/synthetic.py:10-106``) -- same formulas, same
``meshgrid`` 'xy' indexing quirk (the analytic "x" varies along numpy axis 1).

``compute_normal_volume`` is on the hot path: it calls
the sm_100a stencil kernel through the C ABI (``pyvr_cuda_compute_normals``) and
reproduces ``np.gradient`` + ``/(norm + 1e-8)`` of reference
``synthetic.py:109-122`` in float32.  There is no CPU fallback: without the CUDA
library or a GPU it raises.
"""

from __future__ import annotations

import numpy as np

SHAPES = ("sphere", "torus", "double_sphere", "cube", "helix", "random_blob")


def _gauss(d):
    return np.exp(-(d ** 2))


def _random_blob(x, y, z, size):
    from scipy.ndimage import gaussian_filter

    np.random.seed(42)
    off = np.random.uniform(-1, 1, size=3)
    noise = gaussian_filter(np.random.random((size, size, size)).astype(np.float32), sigma=size / 18)
    ramp = ((x + off[0] * 0.5) + 1.5) * ((y + off[1] * 0.5) + 1.2) * ((z + off[2] * 0.5) + 0.8)
    ramp = ramp / np.max(np.abs(ramp))
    vol = np.maximum(0, noise * (0.7 + 0.3 * ramp) - 0.25) * 2.5
    for _ in range(3):
        cx, cy, cz = np.random.uniform(-0.7, 0.7, 3)
        dist = np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2)
        vol += _gauss(dist * 6) * np.random.uniform(0.5, 1.2)
    return np.clip(vol, 0, 1)


def create_sample_volume(size: int = 64, shape: str = "sphere") -> np.ndarray:
    """``(size, size, size) float32`` analytic test volume on ``linspace(-1, 1, size)``."""
    if shape not in SHAPES:
        raise ValueError(
            f"Unknown shape: {shape}. Available shapes: sphere, torus, double_sphere, cube, helix, random_blob")
    axis = np.linspace(-1, 1, size)
    # default 'xy' indexing, as the reference; sparse=True broadcasts the same element-wise
    # arithmetic without materialising three size^3 coordinate arrays (values are bit-identical)
    x, y, z = np.meshgrid(axis, axis, axis, sparse=True)

    if shape == "sphere":
        vol = _gauss(np.sqrt(x * x + y * y + z * z) * 3)
    elif shape == "torus":
        ring = np.sqrt(x * x + y * y) - 0.6
        vol = _gauss(np.sqrt(ring ** 2 + z * z) / 0.3 * 4)
    elif shape == "double_sphere":
        a = _gauss(np.sqrt((x - 0.3) ** 2 + y * y + z * z) * 4)
        b = _gauss(np.sqrt((x + 0.3) ** 2 + y * y + z * z) * 4)
        vol = np.maximum(a, b)
    elif shape == "cube":
        cheb = np.maximum(np.maximum(np.abs(x), np.abs(y)), np.abs(z))
        vol = _gauss((cheb - 0.4) * 8).astype(np.float32)
        vol[cheb > 0.6] = 0
    elif shape == "helix":
        phase = z * 3 * 2 * np.pi
        d = np.sqrt((x - 0.5 * np.cos(phase)) ** 2 + (y - 0.5 * np.sin(phase)) ** 2)
        vol = _gauss(d / 0.15 * 3)
    else:
        vol = _random_blob(x, y, z, size)
    return vol.astype(np.float32)


def compute_normal_volume(volume: np.ndarray, relaxed: bool = False) -> np.ndarray:
    """Normalised central-difference gradient, ``(D, H, W) -> (D, H, W, 3) float32``.

    Runs on ``cuda:0`` through ``pyvr_cuda_compute_normals`` (host buffers in,
    host buffers out).  Component order follows the numpy axes (0, 1, 2).  The result is bit-identical
    to the reference's numpy function; ``relaxed=True`` (not in the reference) trades that for speed:
    quotients within 2 ulp (the path's tolerance is 1e-5 relative).
    """
    from .cuda_renderer import _cabi

    vol = np.ascontiguousarray(volume, dtype=np.float32)
    if vol.ndim != 3:
        raise ValueError(f"Volume data must be 3D, got shape {vol.shape}")
    return _cabi.compute_normals_host(vol, relaxed=relaxed)
