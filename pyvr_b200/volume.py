"""Scalar volume + optional normal volume + world bounds (input container of the path).

Host mirror of the reference's ``Volume`` dataclass (``pyvr/volume/data.py:15-180``):
same fields, defaults (bounds +-0.5), validation messages and helpers.  The one
behavioural difference is where the work happens: :meth:`Volume.compute_normals`
runs the sm_100a gradient stencil (``pyvr_cuda_compute_normals``) instead of
``np.gradient`` -- same values, see ``tests/test_normals_gpu.py``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


def _vec3(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


@dataclass
class Volume:
    data: np.ndarray
    normals: Optional[np.ndarray] = None
    min_bounds: np.ndarray = field(default_factory=lambda: _vec3(-0.5, -0.5, -0.5))
    max_bounds: np.ndarray = field(default_factory=lambda: _vec3(0.5, 0.5, 0.5))
    name: Optional[str] = None

    def __post_init__(self):
        self.validate()

    def validate(self) -> None:
        if not isinstance(self.data, np.ndarray):
            raise ValueError("Volume data must be a numpy array")
        if self.data.ndim != 3:
            raise ValueError(f"Volume data must be 3D, got shape {self.data.shape}")
        if self.normals is not None:
            if not isinstance(self.normals, np.ndarray):
                raise ValueError("Normal volume must be a numpy array")
            want = self.data.shape + (3,)
            if self.normals.shape != want:
                raise ValueError(
                    f"Normal volume must have shape {want}, got {self.normals.shape}")
        for label, b in (("min_bounds", self.min_bounds), ("max_bounds", self.max_bounds)):
            if not isinstance(b, np.ndarray) or b.shape != (3,):
                raise ValueError(f"{label} must be a 3D numpy array")
        if np.any(self.max_bounds <= self.min_bounds):
            raise ValueError("max_bounds must be greater than min_bounds")

    # -- geometry ------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, int, int]:
        return self.data.shape

    @property
    def dimensions(self) -> np.ndarray:
        return self.max_bounds - self.min_bounds

    @property
    def center(self) -> np.ndarray:
        return (self.min_bounds + self.max_bounds) / 2.0

    @property
    def has_normals(self) -> bool:
        return self.normals is not None

    @property
    def voxel_spacing(self) -> np.ndarray:
        return self.dimensions / np.array(self.shape, dtype=np.float32)

    # -- derived volumes -----------------------------------------------------
    def compute_normals(self, method: str = "gradient") -> None:
        if method != "gradient":
            raise ValueError(f"Unsupported method: {method}")
        from .datasets import compute_normal_volume

        self.normals = compute_normal_volume(self.data)

    def normalize(self, method: str = "minmax") -> "Volume":
        if method == "minmax":
            lo, hi = self.data.min(), self.data.max()
            out = np.zeros_like(self.data) if hi - lo < 1e-9 else (self.data - lo) / (hi - lo)
        elif method == "zscore":
            mean, std = self.data.mean(), self.data.std()
            out = np.zeros_like(self.data) if std < 1e-9 else (self.data - mean) / std
        else:
            raise ValueError(f"Unsupported method: {method}")
        return Volume(
            data=out.astype(np.float32),
            normals=None if self.normals is None else self.normals.copy(),
            min_bounds=self.min_bounds.copy(),
            max_bounds=self.max_bounds.copy(),
            name=f"{self.name}_normalized" if self.name else None,
        )

    def copy(self) -> "Volume":
        return Volume(
            data=self.data.copy(),
            normals=None if self.normals is None else self.normals.copy(),
            min_bounds=self.min_bounds.copy(),
            max_bounds=self.max_bounds.copy(),
            name=self.name,
        )

    def __repr__(self) -> str:
        label = f"'{self.name}'" if self.name else "unnamed"
        nrm = "with normals" if self.has_normals else "no normals"
        return (f"Volume({label}, shape={self.shape}, "
                f"bounds=[{self.min_bounds}, {self.max_bounds}], {nrm})")


class VolumeError(Exception):
    """Raised for volume data errors (reference volume/data.py:177)."""
