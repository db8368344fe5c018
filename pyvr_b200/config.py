"""Ray-march quality parameters (host mirror of the reference's ``RenderConfig``).

Mirrors ``pyvr/config.py:11-374`` of the reference: same field names, defaults,
preset values and validation errors, because those fields are the parameter
contract of the march kernel.  Only three of the five fields reach the device
(``step_size``, ``max_steps``, ``reference_step_size`` -- reference
``pyvr/moderngl_renderer/renderer.py:303-309``); the early-termination fields
are carried for API compatibility and used only by the opt-in non-parity mode of
:class:`pyvr_b200.cuda_renderer.VolumeRenderer`.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass, replace

# (step_size, max_steps, early_ray_termination, opacity_threshold); reference config.py:96-227
_PRESETS = {
    "preview": (0.05, 50, True, 0.80),
    "fast": (0.02, 100, True, 0.90),
    "balanced": (0.01, 500, True, 0.95),
    "high_quality": (0.005, 1000, True, 0.98),
    "ultra_quality": (0.001, 2000, False, 1.0),
}

# Diagonal of the unit cube used by the reference's sample estimator (config.py:330).
_UNIT_DIAGONAL = 1.732


@dataclass
class RenderConfig:
    step_size: float = 0.01
    max_steps: int = 500
    early_ray_termination: bool = True
    opacity_threshold: float = 0.95
    reference_step_size: float = 0.01

    def __post_init__(self):
        self.validate()

    def validate(self) -> None:
        if self.step_size <= 0:
            raise ValueError("step_size must be positive")
        if self.step_size > 1.0:
            warnings.warn(
                f"step_size {self.step_size} is very large and may produce poor quality",
                UserWarning,
            )
        if self.max_steps < 1:
            raise ValueError("max_steps must be at least 1")
        if self.max_steps > 10000:
            warnings.warn(
                f"max_steps {self.max_steps} is very large and may be slow", UserWarning
            )
        if not (0.0 <= self.opacity_threshold <= 1.0):
            raise ValueError("opacity_threshold must be between 0.0 and 1.0")

    # -- presets -----------------------------------------------------------
    @classmethod
    def _preset(cls, name: str) -> "RenderConfig":
        step, steps, ert, thr = _PRESETS[name]
        return cls(step_size=step, max_steps=steps, early_ray_termination=ert,
                   opacity_threshold=thr)

    @classmethod
    def preview(cls) -> "RenderConfig":
        return cls._preset("preview")

    @classmethod
    def fast(cls) -> "RenderConfig":
        return cls._preset("fast")

    @classmethod
    def balanced(cls) -> "RenderConfig":
        return cls._preset("balanced")

    @classmethod
    def high_quality(cls) -> "RenderConfig":
        return cls._preset("high_quality")

    @classmethod
    def ultra_quality(cls) -> "RenderConfig":
        return cls._preset("ultra_quality")

    @classmethod
    def custom(cls, step_size: float, max_steps: int, early_ray_termination: bool = True,
               opacity_threshold: float = 0.95) -> "RenderConfig":
        return cls(step_size=step_size, max_steps=max_steps,
                   early_ray_termination=early_ray_termination,
                   opacity_threshold=opacity_threshold)

    # -- derived copies ----------------------------------------------------
    def copy(self) -> "RenderConfig":
        return replace(self)

    def with_step_size(self, step_size: float) -> "RenderConfig":
        return replace(self, step_size=step_size)

    def with_max_steps(self, max_steps: int) -> "RenderConfig":
        return replace(self, max_steps=max_steps)

    # -- estimators ----------------------------------------------------------
    def estimate_samples_per_ray(self) -> int:
        return min(int(_UNIT_DIAGONAL / self.step_size), self.max_steps)

    def estimate_render_time_relative(self) -> float:
        return self.estimate_samples_per_ray() / RenderConfig.balanced().estimate_samples_per_ray()

    def __repr__(self) -> str:
        return (f"RenderConfig(step_size={self.step_size}, max_steps={self.max_steps}, "
                f"early_termination={self.early_ray_termination}, "
                f"~{self.estimate_samples_per_ray()} samples/ray)")


class RenderConfigError(Exception):
    """Raised for rendering-configuration errors (reference config.py:371)."""
