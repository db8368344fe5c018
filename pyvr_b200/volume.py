"""Scalar volume + optional normal volume + world bounds (input container of the path).

Host mirror of the reference's ``Volume`` dataclass (``pyvr/volume/data.py:15-180``): same
fields, defaults (bounds +-0.5), validation messages and helpers, so reference scripts and the
reference's own tests run against it unchanged.  The one behavioural difference is where the work
happens: :meth:`Volume.compute_normals` runs the sm_100a gradient stencil
(``pyvr_cuda_compute_normals``) instead of ``np.gradient`` -- same values, see
``tests/test_normals_gpu.py``.
"""

from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Callable, Dict, Optional, Tuple

import numpy as np

_HALF = np.float32(0.5)


def _corner(sign: float) -> Callable[[], np.ndarray]:
    return lambda: np.full(3, sign * _HALF, dtype=np.float32)


def _problems(vol: "Volume"):
    """Yield the message of every violated invariant, in the order the reference checks them."""
    data, normals = vol.data, vol.normals
    if not isinstance(data, np.ndarray):
        yield "Volume data must be a numpy array"
        return
    if data.ndim != 3:
        yield f"Volume data must be 3D, got shape {data.shape}"
        return
    if normals is not None:
        if not isinstance(normals, np.ndarray):
            yield "Normal volume must be a numpy array"
            return
        if normals.shape != (*data.shape, 3):
            yield f"Normal volume must have shape {(*data.shape, 3)}, got {normals.shape}"
            return
    for label in ("min_bounds", "max_bounds"):
        corner = getattr(vol, label)
        if not (isinstance(corner, np.ndarray) and corner.shape == (3,)):
            yield f"{label} must be a 3D numpy array"
            return
    if np.any(vol.max_bounds <= vol.min_bounds):
        yield "max_bounds must be greater than min_bounds"


# normalisation rules: data -> (offset, scale); a degenerate scale maps everything to zero
_NORMALISERS: Dict[str, Callable[[np.ndarray], Tuple[float, float]]] = {
    "minmax": lambda d: (d.min(), d.max() - d.min()),
    "zscore": lambda d: (d.mean(), d.std()),
}


@dataclass
class Volume:
    data: np.ndarray
    normals: Optional[np.ndarray] = None
    min_bounds: np.ndarray = field(default_factory=_corner(-1.0))
    max_bounds: np.ndarray = field(default_factory=_corner(+1.0))
    name: Optional[str] = None

    def __post_init__(self):
        self.validate()

    def validate(self) -> None:
        for message in _problems(self):
            raise ValueError(message)

    # -- geometry ------------------------------------------------------------
    shape = property(lambda self: self.data.shape, doc="(D, H, W) of the scalar array")
    dimensions = property(lambda self: self.max_bounds - self.min_bounds, doc="world-space extent of the box")
    center = property(lambda self: (self.min_bounds + self.max_bounds) / 2.0, doc="world-space centre of the box")
    has_normals = property(lambda self: self.normals is not None, doc="whether a normal volume is attached")
    voxel_spacing = property(lambda self: self.dimensions / np.asarray(self.shape, dtype=np.float32),
                             doc="world-space size of one voxel")

    # -- derived volumes -----------------------------------------------------
    def compute_normals(self, method: str = "gradient") -> None:
        if method != "gradient":
            raise ValueError(f"Unsupported method: {method}")
        from .datasets import compute_normal_volume

        self.normals = compute_normal_volume(self.data)

    def _clone(self, data: np.ndarray, name: Optional[str]) -> "Volume":
        return replace(self, data=data, name=name, min_bounds=self.min_bounds.copy(), max_bounds=self.max_bounds.copy(),
                       normals=self.normals.copy() if self.has_normals else None)

    def normalize(self, method: str = "minmax") -> "Volume":
        rule = _NORMALISERS.get(method)
        if rule is None:
            raise ValueError(f"Unsupported method: {method}")
        offset, scale = rule(self.data)
        scaled = np.zeros_like(self.data) if scale < 1e-9 else (self.data - offset) / scale
        return self._clone(scaled.astype(np.float32), f"{self.name}_normalized" if self.name else None)

    def copy(self) -> "Volume":
        return self._clone(self.data.copy(), self.name)

    def __repr__(self) -> str:
        label = f"'{self.name}'" if self.name else "unnamed"
        tail = "with normals" if self.has_normals else "no normals"
        return f"Volume({label}, shape={self.shape}, bounds=[{self.min_bounds}, {self.max_bounds}], {tail})"


class VolumeError(Exception):
    """Raised for volume data errors (reference volume/data.py:177)."""
