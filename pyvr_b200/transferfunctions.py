"""Colour / opacity transfer functions and the RGBA LUT the march kernel samples.

Host mirror of ``pyvr/transferfunctions/{base,color,opacity}.py`` of the
reference.  The only product of this module that reaches the device is the
``(size, 4) float32`` RGBA table built by :func:`build_rgba_lut`, which restates
``ModernGLManager.create_rgba_transfer_function_texture``
(``pyvr/moderngl_renderer/manager.py:163-175``): ``np.interp`` of the control
points over ``linspace(0, 1, size)``, RGB from the colour TF, A from the opacity TF.

Colormap names: the reference samples matplotlib (``color.py:126-143``).  When
matplotlib is not installed this module falls back to vendored 8-bit tables
(``_colormap_tables.py``, from OpenCV; within 0.5/255 of matplotlib's floats).
"""

from __future__ import annotations

import base64
from abc import ABC, abstractmethod
from typing import List, Optional, Sequence, Tuple

import numpy as np


class TransferFunctionError(Exception):
    pass


class InvalidControlPointError(TransferFunctionError):
    pass


def validate_control_points_format(control_points: Sequence, expected_value_length: int) -> None:
    """Shape/type check of ``[(scalar, value), ...]`` (reference base.py:113-161)."""
    if not control_points:
        raise InvalidControlPointError("Control points cannot be empty")
    for i, point in enumerate(control_points):
        if not isinstance(point, (tuple, list)) or len(point) != 2:
            raise InvalidControlPointError(
                f"Control point {i} must be a (scalar, value) tuple, got {point}")
        scalar, value = point
        if not isinstance(scalar, (int, float)):
            raise InvalidControlPointError(
                f"Control point {i} scalar must be numeric, got {type(scalar)}")
        if expected_value_length == 1:
            if not isinstance(value, (int, float)):
                raise InvalidControlPointError(
                    f"Control point {i} value must be numeric, got {type(value)}")
            continue
        if (not isinstance(value, (tuple, list, np.ndarray))
                or len(value) != expected_value_length):
            raise InvalidControlPointError(
                f"Control point {i} value must be a {expected_value_length}-element sequence, got {value}")
        for j, component in enumerate(value):
            if not isinstance(component, (int, float)):
                raise InvalidControlPointError(
                    f"Control point {i} value component {j} must be numeric, got {type(component)}")


class BaseTransferFunction(ABC):
    def __init__(self, control_points: Optional[List[Tuple]] = None, lut_size: int = 256):
        if control_points is None:
            control_points = self._get_default_control_points()
        self.control_points = sorted(control_points, key=lambda p: p[0])
        self.lut_size = lut_size
        self._validate_control_points()

    @abstractmethod
    def _get_default_control_points(self) -> List[Tuple]: ...

    @abstractmethod
    def _validate_control_points(self) -> None: ...

    @abstractmethod
    def to_lut(self, size: Optional[int] = None) -> np.ndarray: ...

    def __call__(self, size: Optional[int] = None) -> np.ndarray:
        return self.to_lut(size)

    def _validate_scalar_range(self, min_val: float = 0.0, max_val: float = 1.0) -> None:
        for scalar, _ in self.control_points:
            if not (min_val <= scalar <= max_val):
                raise ValueError(
                    f"Control point scalar {scalar} outside valid range [{min_val}, {max_val}]")

    def _get_scalar_values(self) -> np.ndarray:
        return np.array([p[0] for p in self.control_points])

    def _get_mapped_values(self) -> np.ndarray:
        return np.array([p[1] for p in self.control_points])


def _colormap_rgb(name: str, x: np.ndarray) -> np.ndarray:
    """RGB rows for positions ``x`` in [0, 1] of a named colormap."""
    try:
        import matplotlib

        return np.asarray(matplotlib.colormaps.get_cmap(name)(x))[:, :3]
    except ImportError:
        pass
    from ._colormap_tables import TABLES_B64

    if name not in TABLES_B64:
        raise ValueError(
            f"Unknown colormap name '{name}': matplotlib is not installed and the vendored "
            f"tables only hold {sorted(TABLES_B64)}")
    table = np.frombuffer(base64.b64decode("".join(TABLES_B64[name])), np.uint8)
    table = table.reshape(256, 3).astype(np.float64) / 255.0
    # matplotlib's ListedColormap lookup: index = floor(x * N) clipped to N-1
    idx = np.clip((x * 256).astype(np.int64), 0, 255)
    return table[idx]


class ColorTransferFunction(BaseTransferFunction):
    """scalar -> RGB, piecewise linear over control points (reference color.py:21-248)."""

    def _get_default_control_points(self):
        return [(0.0, (0.0, 0.0, 0.0)), (1.0, (1.0, 1.0, 1.0))]

    def _validate_control_points(self) -> None:
        validate_control_points_format(self.control_points, expected_value_length=3)
        self._validate_scalar_range(0.0, 1.0)
        for i, (_, rgb) in enumerate(self.control_points):
            for j, c in enumerate(rgb):
                if not (0.0 <= c <= 1.0):
                    raise InvalidControlPointError(
                        f"Control point {i} RGB component {j} value {c} outside valid range [0.0, 1.0]")

    @classmethod
    def grayscale(cls, lut_size: int = 256) -> "ColorTransferFunction":
        return cls([(0.0, (0.0, 0.0, 0.0)), (1.0, (1.0, 1.0, 1.0))], lut_size=lut_size)

    @classmethod
    def single_color(cls, color, lut_size: int = 256) -> "ColorTransferFunction":
        return cls([(0.0, color), (1.0, color)], lut_size=lut_size)

    @classmethod
    def two_color_ramp(cls, color1, color2, lut_size: int = 256) -> "ColorTransferFunction":
        return cls([(0.0, color1), (1.0, color2)], lut_size=lut_size)

    @classmethod
    def from_colormap(cls, colormap_name: str, value_range: Tuple[float, float] = (0.0, 1.0),
                      lut_size: int = 256) -> "ColorTransferFunction":
        x = np.linspace(0, 1, lut_size)
        colors = _colormap_rgb(colormap_name, x)
        lo, hi = value_range
        xs = lo + x * (hi - lo)
        points = [(float(xi), tuple(map(float, rgb))) for xi, rgb in zip(xs, colors)]
        return cls(points, lut_size=lut_size)

    def _interp(self, x: np.ndarray) -> np.ndarray:
        scalars, colors = zip(*self.control_points)
        colors = np.array(colors)
        out = np.empty((x.size, 3), dtype=np.float32)
        for c in range(3):
            out[:, c] = np.interp(x, scalars, colors[:, c])
        return out

    def to_lut(self, size: Optional[int] = None) -> np.ndarray:
        return self._interp(np.linspace(0, 1, size or self.lut_size))

    def apply_to_array(self, scalar_array: np.ndarray) -> np.ndarray:
        return self._interp(scalar_array.ravel()).reshape(scalar_array.shape + (3,))

    def get_color_at(self, scalar: float):
        return tuple(self._interp(np.array([scalar]))[0].tolist())

    def __repr__(self) -> str:
        return (f"ColorTransferFunction({len(self.control_points)} control points, "
                f"lut_size={self.lut_size})")


class OpacityTransferFunction(BaseTransferFunction):
    """scalar -> opacity at the reference step size (reference opacity.py:21-214)."""

    def _get_default_control_points(self):
        return [(0.0, 0.0), (1.0, 1.0)]

    def _validate_control_points(self) -> None:
        validate_control_points_format(self.control_points, expected_value_length=1)
        self._validate_scalar_range(0.0, 1.0)
        for i, (_, opacity) in enumerate(self.control_points):
            if not (0.0 <= opacity <= 1.0):
                raise InvalidControlPointError(
                    f"Control point {i} opacity {opacity} outside valid range [0.0, 1.0]")

    @classmethod
    def linear(cls, low: float = 0.0, high: float = 1.0, lut_size: int = 256):
        return cls([(0.0, low), (1.0, high)], lut_size=lut_size)

    @classmethod
    def one_step(cls, step: float = 0.5, low: float = 0.0, high: float = 1.0, lut_size: int = 256):
        return cls([(0.0, low), (step, low), (step + 1e-12, high), (1.0, high)], lut_size=lut_size)

    @classmethod
    def peaks(cls, peaks: List[float], opacity: float = 1.0, eps: float = 0.02,
              lut_size: int = 256, base: float = 0.0):
        if not peaks:
            raise ValueError("At least one peak position must be specified")
        for p in peaks:
            if not (0.0 <= p <= 1.0):
                raise ValueError(f"Peak position {p} must be between 0 and 1")
        points = [(0.0, base)]
        for p in sorted(peaks):
            left, right = max(0.0, p - eps), min(1.0, p + eps)
            if left > points[-1][0]:
                points.append((left, base))
            points.append((p, opacity))
            if right > p:
                points.append((right, base))
        if points[-1][0] < 1.0:
            points.append((1.0, base))
        return cls(points, lut_size=lut_size)

    def to_lut(self, size: Optional[int] = None) -> np.ndarray:
        scalars, opacities = zip(*self.control_points)
        x = np.linspace(0, 1, size or self.lut_size)
        return np.interp(x, scalars, opacities).astype(np.float32)

    def apply_to_array(self, scalar_array: np.ndarray) -> np.ndarray:
        scalars, opacities = zip(*self.control_points)
        flat = np.interp(scalar_array.ravel(), scalars, opacities)
        return flat.reshape(scalar_array.shape).astype(np.float32)

    def __repr__(self) -> str:
        return (f"OpacityTransferFunction(control_points={self.control_points}, "
                f"lut_size={self.lut_size})")


def build_rgba_lut(color_tf, opacity_tf, size: Optional[int] = None) -> np.ndarray:
    """The ``(size, 4) float32`` RGBA table uploaded to the device.

    Restates reference ``manager.py:163-175``: ``size`` defaults to the larger of
    the two ``lut_size`` attributes; RGB = ``color_tf.to_lut(size)``, A =
    ``opacity_tf.to_lut(size)``.  Works with the reference's own TF objects too
    (only ``lut_size`` and ``to_lut`` are used).
    """
    if size is None:
        size = max(color_tf.lut_size, opacity_tf.lut_size)
    lut = np.empty((size, 4), dtype=np.float32)
    lut[:, :3] = color_tf.to_lut(size)
    lut[:, 3] = opacity_tf.to_lut(size)
    return lut
