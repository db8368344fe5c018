"""pyvr_b200 -- B200-native (sm_100a) backend for PyVR's volume ray-march path.

Host-side mirror of the reference's object model (``Volume``, ``Camera``, ``Light``,
``RenderConfig``, transfer functions) plus ``pyvr_b200.cuda_renderer.VolumeRenderer``, the
drop-in for ``pyvr.moderngl_renderer.VolumeRenderer``.  All compute runs in hand-written CUDA
behind the C ABI declared in ``include/pyvr_cuda.h``; there is no CPU fallback.
"""

from .camera import Camera, CameraError, get_camera_pos, get_camera_pos_from_params
from .config import RenderConfig, RenderConfigError
from .datasets import compute_normal_volume, create_sample_volume
from .lighting import Light, LightError
from .transferfunctions import (ColorTransferFunction, OpacityTransferFunction,
                                build_rgba_lut)
from .volume import Volume, VolumeError

__version__ = "0.1.0"

__all__ = [
    "Camera", "CameraError", "get_camera_pos", "get_camera_pos_from_params",
    "RenderConfig", "RenderConfigError", "compute_normal_volume", "create_sample_volume",
    "Light", "LightError", "ColorTransferFunction", "OpacityTransferFunction",
    "build_rgba_lut", "Volume", "VolumeError",
]
