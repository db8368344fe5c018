"""Light description consumed by the march kernel's Lambert term.

Host mirror of the reference's ``Light`` (``pyvr/lighting/light.py:15-431``).
The device uses four things from it (reference ``renderer.py:311-316`` and
``volume.frag.glsl:107-110``): ``ambient_intensity``, ``diffuse_intensity`` and
the direction of travel ``normalize(target - position)``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


def _f32(values) -> np.ndarray:
    return np.array(values, dtype=np.float32)


@dataclass
class Light:
    position: np.ndarray = field(default_factory=lambda: _f32([1.0, 1.0, 1.0]))
    target: np.ndarray = field(default_factory=lambda: _f32([0.0, 0.0, 0.0]))
    ambient_intensity: float = 0.2
    diffuse_intensity: float = 0.8

    _is_linked: bool = field(default=False, init=False, repr=False)
    _camera_offsets: Optional[dict] = field(default=None, init=False, repr=False)

    def __post_init__(self):
        self.validate()

    def validate(self) -> None:
        if not isinstance(self.position, np.ndarray) or self.position.shape != (3,):
            raise ValueError("position must be a 3D numpy array")
        if not isinstance(self.target, np.ndarray) or self.target.shape != (3,):
            raise ValueError("target must be a 3D numpy array")
        if not (0.0 <= self.ambient_intensity <= 1.0):
            raise ValueError("ambient_intensity must be between 0.0 and 1.0")
        if not (0.0 <= self.diffuse_intensity <= 1.0):
            raise ValueError("diffuse_intensity must be between 0.0 and 1.0")

    # -- presets (reference light.py:67-260) -----------------------------------
    @classmethod
    def directional(cls, direction, ambient: float = 0.2, diffuse: float = 0.8,
                    distance: float = 10.0) -> "Light":
        d = _f32(direction)
        d = d / np.linalg.norm(d)
        return cls(position=-d * distance, target=_f32([0, 0, 0]),
                   ambient_intensity=ambient, diffuse_intensity=diffuse)

    @classmethod
    def point_light(cls, position, target=None, ambient: float = 0.2,
                    diffuse: float = 0.8) -> "Light":
        return cls(position=_f32(position),
                   target=_f32([0, 0, 0]) if target is None else _f32(target),
                   ambient_intensity=ambient, diffuse_intensity=diffuse)

    @classmethod
    def default(cls) -> "Light":
        return cls()

    @classmethod
    def ambient_only(cls, intensity: float = 0.5) -> "Light":
        return cls(position=_f32([0, 0, 0]), target=_f32([0, 0, 0]),
                   ambient_intensity=intensity, diffuse_intensity=0.0)

    @classmethod
    def camera_linked(cls, azimuth_offset: float = 0.0, elevation_offset: float = 0.0,
                      distance_offset: float = 0.0, ambient: float = 0.2,
                      diffuse: float = 0.8) -> "Light":
        light = cls(ambient_intensity=ambient, diffuse_intensity=diffuse)
        light.link_to_camera(azimuth_offset, elevation_offset, distance_offset)
        return light

    # -- queries ---------------------------------------------------------------
    def get_direction(self) -> np.ndarray:
        d = self.target - self.position
        n = np.linalg.norm(d)
        if n < 1e-9:
            return _f32([0.0, 0.0, -1.0])
        return d / n

    def copy(self) -> "Light":
        twin = Light(position=self.position.copy(), target=self.target.copy(),
                     ambient_intensity=self.ambient_intensity,
                     diffuse_intensity=self.diffuse_intensity)
        twin._is_linked = self._is_linked
        if self._camera_offsets is not None:
            twin._camera_offsets = dict(self._camera_offsets)
        return twin

    # -- camera linking ("headlight", reference light.py:301-414) -----------------
    @property
    def is_linked(self) -> bool:
        return self._is_linked

    def link_to_camera(self, azimuth_offset: float = 0.0, elevation_offset: float = 0.0,
                       distance_offset: float = 0.0) -> "Light":
        self._is_linked = True
        self._camera_offsets = {"azimuth": azimuth_offset, "elevation": elevation_offset,
                                "distance": distance_offset}
        return self

    def unlink_from_camera(self) -> "Light":
        self._is_linked = False
        self._camera_offsets = None
        return self

    def update_from_camera(self, camera) -> None:
        if not self._is_linked:
            raise ValueError("Light is not linked to camera. Call link_to_camera() first.")
        if self._camera_offsets is None:
            raise ValueError("Camera offsets not set. Call link_to_camera() first.")
        if not (hasattr(camera, "get_camera_vectors") and hasattr(camera, "target")):
            raise ValueError("camera must be a Camera instance")
        position, _ = camera.get_camera_vectors()
        self.position = position.copy()
        self.target = camera.target.copy()

    def get_offsets(self) -> Optional[dict]:
        return dict(self._camera_offsets) if self._camera_offsets else None

    def __repr__(self) -> str:
        return (f"Light(position={self.position}, target={self.target}, "
                f"ambient={self.ambient_intensity:.2f}, diffuse={self.diffuse_intensity:.2f})")


class LightError(Exception):
    """Raised for lighting configuration errors (reference light.py:428)."""
