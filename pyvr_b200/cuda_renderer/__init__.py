"""``pyvr.cuda_renderer`` backend: ``from pyvr_b200.cuda_renderer import VolumeRenderer``.

Mirrors the export list of the reference's ``pyvr/moderngl_renderer/__init__.py:7-27`` so that
reference scripts switch backend by changing the import.
"""

from ..camera import Camera
from ..datasets import compute_normal_volume, create_sample_volume
from ..transferfunctions import ColorTransferFunction, OpacityTransferFunction
from .renderer import CudaVolumeRenderer, VolumeRenderer

__all__ = [
    "ColorTransferFunction", "OpacityTransferFunction", "CudaVolumeRenderer", "VolumeRenderer",
    "Camera", "create_sample_volume", "compute_normal_volume",
]
