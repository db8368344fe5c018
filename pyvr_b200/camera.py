"""Spherical camera -> (position, up) -> view / projection matrices.

Host mirror of the parts of the reference's camera package the render path
consumes: ``Camera`` (``pyvr/camera/camera.py:16-380``) and ``get_camera_pos``
(``pyvr/camera/control.py:238-351``).  The interactive helpers of the reference
(``CameraPath``, ``CameraController``, trackball) are GUI-side callers and are
out of scope.

The reference composes three scipy ``Rotation.from_rotvec`` rotations; here the
same rotations are written as explicit Rodrigues matrices in float64, keeping
the reference's float32 preparation of the input vectors so the resulting
position/up agree with it to ~1e-15 (pinned by ``tests/golden/camera.json``).
"""

from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np


class CameraError(Exception):
    """Raised for invalid camera parameters (reference camera.py:383)."""


def validate_camera_angles(azimuth: float, elevation: float, roll: float) -> None:
    for name, value in (("azimuth", azimuth), ("elevation", elevation), ("roll", roll)):
        if not isinstance(value, (int, float)):
            raise CameraError(f"{name} must be numeric, got {type(value)}")
        if not np.isfinite(value):
            raise CameraError(f"{name} must be finite, got {value}")


def degrees_to_radians(**kwargs) -> Dict[str, float]:
    return {k: np.radians(v) for k, v in kwargs.items()}


def radians_to_degrees(**kwargs) -> Dict[str, float]:
    return {k: np.degrees(v) for k, v in kwargs.items()}


def _rotvec_matrix(rotvec: np.ndarray) -> np.ndarray:
    """Rotation matrix of a rotation vector (axis * angle), float64 Rodrigues."""
    v = np.asarray(rotvec, dtype=np.float64)
    angle = float(np.linalg.norm(v))
    if angle < 1e-300:
        return np.eye(3)
    k = v / angle
    K = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * (K @ K)


def get_camera_pos(
    target: np.ndarray,
    azimuth: float,
    elevation: float,
    roll: float,
    distance: float,
    init_pos: Optional[np.ndarray] = None,
    init_up: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, np.ndarray]:
    """Camera position and up vector from spherical parameters.

    Follows reference ``control.py:293-351``: the start offset is ``init_pos``
    rescaled to ``distance``; azimuth turns about ``init_up``, elevation about
    ``init_up x view_dir``, roll about the view direction, composed as
    ``R_az . R_el . R_roll`` and applied to the offset and to ``init_up``.
    """
    validate_camera_angles(azimuth, elevation, roll)
    if not isinstance(target, np.ndarray) or target.shape != (3,):
        raise ValueError("target must be a 3D numpy array")
    if not isinstance(distance, (int, float)) or distance <= 0:
        raise ValueError("distance must be positive")
    if init_pos is None:
        init_pos = np.array([0, 0, distance], dtype=np.float32)
    if init_up is None:
        init_up = np.array([0, 1, 0], dtype=np.float32)
    if not isinstance(init_pos, np.ndarray) or init_pos.shape != (3,):
        raise ValueError("init_pos must be a 3D numpy array")
    if not isinstance(init_up, np.ndarray) or init_up.shape != (3,):
        raise ValueError("init_up must be a 3D numpy array")

    # The reference does this preparation in float32 (control.py:312-325).
    target = target.astype(np.float32)
    init_pos = init_pos.astype(np.float32)
    init_up = init_up.astype(np.float32)
    offset = init_pos - target
    length = np.linalg.norm(offset)
    if length == 0:
        raise ValueError("init_pos must not be the zero vector (relative to target)")
    offset = offset / length * np.float32(distance)
    view_dir = -offset / np.linalg.norm(offset)

    elev_axis = np.cross(init_up, view_dir)
    elev_len = np.linalg.norm(elev_axis)
    if elev_len < 1e-6:  # up parallel to the view direction: pick any perpendicular axis
        helper = np.array([1, 0, 0]) if abs(init_up[0]) < 0.9 else np.array([0, 0, 1])
        elev_axis = np.cross(init_up, helper)
        elev_len = np.linalg.norm(elev_axis)
    elev_axis = elev_axis / elev_len

    rot = (_rotvec_matrix(azimuth * init_up)
           @ _rotvec_matrix(elevation * elev_axis)
           @ _rotvec_matrix(roll * view_dir))
    position = rot @ offset.astype(np.float64) + target
    up = rot @ init_up.astype(np.float64)
    return position, up


def get_camera_pos_from_params(params: "Camera") -> Tuple[np.ndarray, np.ndarray]:
    return get_camera_pos(
        target=params.target, azimuth=params.azimuth, elevation=params.elevation,
        roll=params.roll, distance=params.distance,
        init_pos=params.init_pos, init_up=params.init_up,
    )


def _origin() -> np.ndarray:
    return np.array([0.0, 0.0, 0.0], dtype=np.float32)


@dataclass
class Camera:
    target: np.ndarray = field(default_factory=_origin)
    azimuth: float = 0.0
    elevation: float = 0.0
    roll: float = 0.0
    distance: float = 3.0
    init_pos: np.ndarray = field(default_factory=lambda: np.array([1.0, 0.0, 0.0], dtype=np.float32))
    init_up: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 1.0], dtype=np.float32))
    fov: float = np.pi / 4
    near_plane: float = 0.1
    far_plane: float = 100.0

    def __post_init__(self):
        self.validate()

    def validate(self) -> None:
        if not isinstance(self.target, np.ndarray) or self.target.shape != (3,):
            raise ValueError("target must be a 3D numpy array")
        for name in ("azimuth", "elevation", "roll"):
            value = getattr(self, name)
            if not isinstance(value, (int, float)):
                raise ValueError(f"{name} must be numeric")
            if abs(value) > 4 * np.pi:
                print(f"Warning: {name} = {value:.3f} rad ({np.degrees(value):.1f}°) is unusually large")
        if not isinstance(self.distance, (int, float)) or self.distance <= 0:
            raise ValueError("distance must be positive")
        if not isinstance(self.init_pos, np.ndarray) or self.init_pos.shape != (3,):
            raise ValueError("init_pos must be a 3D numpy array")
        if not isinstance(self.init_up, np.ndarray) or self.init_up.shape != (3,):
            raise ValueError("init_up must be a 3D numpy array")
        if np.linalg.norm(self.init_pos - self.target) < 1e-9:
            raise ValueError("init_pos must not be at the same location as target")
        if np.linalg.norm(self.init_up) < 1e-9:
            raise ValueError("init_up must not be the zero vector")
        if self.fov <= 0 or self.fov >= np.pi:
            raise ValueError("fov must be between 0 and π radians")
        if self.near_plane <= 0:
            raise ValueError("near_plane must be positive")
        if self.far_plane <= self.near_plane:
            raise ValueError("far_plane must be greater than near_plane")

    # -- constructors (reference camera.py:113-194) ---------------------------
    @classmethod
    def from_spherical(cls, target, azimuth, elevation, roll, distance, **kwargs) -> "Camera":
        return cls(target=target, azimuth=azimuth, elevation=elevation, roll=roll,
                   distance=distance, **kwargs)

    @classmethod
    def _view(cls, target, distance, azimuth, elevation) -> "Camera":
        return cls(target=_origin() if target is None else target, azimuth=azimuth,
                   elevation=elevation, roll=0.0, distance=distance)

    @classmethod
    def front_view(cls, target=None, distance: float = 3.0) -> "Camera":
        return cls._view(target, distance, 0.0, 0.0)

    @classmethod
    def side_view(cls, target=None, distance: float = 3.0) -> "Camera":
        return cls._view(target, distance, np.pi / 2, 0.0)

    @classmethod
    def top_view(cls, target=None, distance: float = 3.0) -> "Camera":
        return cls._view(target, distance, 0.0, np.pi / 2)

    @classmethod
    def isometric_view(cls, target=None, distance: float = 3.0) -> "Camera":
        return cls._view(target, distance, np.pi / 4, np.pi / 6)

    # -- serialisation (reference camera.py:196-267) -----------------------------
    def to_dict(self) -> Dict[str, Any]:
        return {
            "target": self.target.tolist(),
            "azimuth": float(self.azimuth), "elevation": float(self.elevation),
            "roll": float(self.roll), "distance": float(self.distance),
            "init_pos": self.init_pos.tolist(), "init_up": self.init_up.tolist(),
            "fov": float(self.fov), "near_plane": float(self.near_plane),
            "far_plane": float(self.far_plane),
        }

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "Camera":
        return cls(
            target=np.array(data["target"], dtype=np.float32),
            azimuth=data["azimuth"], elevation=data["elevation"], roll=data["roll"],
            distance=data["distance"],
            init_pos=np.array(data["init_pos"], dtype=np.float32),
            init_up=np.array(data["init_up"], dtype=np.float32),
            fov=data.get("fov", np.pi / 4),
            near_plane=data.get("near_plane", 0.1), far_plane=data.get("far_plane", 100.0),
        )

    def save_to_file(self, filepath: Union[str, Path]) -> None:
        with open(filepath, "w") as f:
            json.dump(self.to_dict(), f, indent=2)

    @classmethod
    def load_from_file(cls, filepath: Union[str, Path]) -> "Camera":
        with open(filepath, "r") as f:
            return cls.from_dict(json.load(f))

    def copy(self) -> "Camera":
        return Camera.from_dict(self.to_dict())

    # -- what the renderer consumes --------------------------------------------
    def get_camera_vectors(self) -> Tuple[np.ndarray, np.ndarray]:
        return get_camera_pos_from_params(self)

    def get_view_matrix(self) -> np.ndarray:
        """Look-at matrix laid out so its row-major bytes are GL column-major
        (reference camera.py:305-334): rows 0-2 hold the basis components,
        row 3 the translation."""
        position, up = self.get_camera_vectors()
        forward = self.target - position
        forward = forward / np.linalg.norm(forward)
        right = np.cross(forward, up)
        right = right / np.linalg.norm(right)
        true_up = np.cross(right, forward)
        m = np.zeros((4, 4), dtype=np.float64)
        m[:3, 0] = right
        m[:3, 1] = true_up
        m[:3, 2] = -forward
        m[3, :3] = (-np.dot(right, position), -np.dot(true_up, position), np.dot(forward, position))
        m[3, 3] = 1.0
        return m.astype(np.float32)

    def get_projection_matrix(self, aspect_ratio: float) -> np.ndarray:
        """Perspective matrix in mathematical (row-major) layout, as the reference
        builds it (camera.py:354-369).  NB: unlike the view matrix it is *not*
        pre-transposed for GL; the render path depends only on its x/y scale
        entries, which the transpose leaves in place."""
        f = 1.0 / np.tan(self.fov / 2.0)
        n, fa = self.near_plane, self.far_plane
        m = np.zeros((4, 4), dtype=np.float64)
        m[0, 0] = f / aspect_ratio
        m[1, 1] = f
        m[2, 2] = (fa + n) / (n - fa)
        m[2, 3] = (2 * fa * n) / (n - fa)
        m[3, 2] = -1.0
        return m.astype(np.float32)

    def __repr__(self) -> str:
        return (f"Camera(target={self.target}, "
                f"azimuth={self.azimuth:.3f} rad ({np.degrees(self.azimuth):.1f}°), "
                f"elevation={self.elevation:.3f} rad ({np.degrees(self.elevation):.1f}°), "
                f"roll={self.roll:.3f} rad ({np.degrees(self.roll):.1f}°), "
                f"distance={self.distance:.2f})")
