"""``VolumeRenderer`` -- the drop-in for ``pyvr.moderngl_renderer.VolumeRenderer``.

Same constructor, methods, defaults and error messages as the reference class
(``pyvr/moderngl_renderer/renderer.py:19-320``); the OpenGL resource manager and the GLSL
shader underneath are replaced by the sm_100a library behind ``include/pyvr_cuda.h``.
``render()`` returns the same ``width*height*4`` RGBA8 bytes, bottom row first, holding the
reference's blended ``(C*A, A*A)`` values.

It accepts this package's host classes and, duck-typed, the reference's own ``pyvr`` objects, so a
reference script switches backend by changing one import.  Extras that have no reference
counterpart (``render_batch``, ``render_accum``, ``stats``) are additive.
"""

from __future__ import annotations

import ctypes
from typing import Iterable, Optional

import numpy as np

from ..camera import Camera
from ..config import RenderConfig
from ..lighting import Light
from ..transferfunctions import build_rgba_lut
from ..volume import Volume
from . import _cabi
from .manager import CudaManager

# The shader's stop test is a literal (volume.frag.glsl:87); RenderConfig.opacity_threshold never
# reaches it (renderer.py:303-309).
REFERENCE_TERMINATION_ALPHA = 0.99


def _is_a(obj, own_cls, name: str) -> bool:
    """True for this package's class or the reference's class of the same name."""
    if isinstance(obj, own_cls):
        return True
    t = type(obj)
    return t.__name__ == name and t.__module__.split(".")[0] == "pyvr"


class CudaVolumeRenderer:
    def __init__(self, width=512, height=512, config=None, light=None, *, device: int = 0,
                 texel_format: str = "f32", strict: bool = False,
                 empty_space_skipping: bool = True, honor_config_termination: bool = False,
                 hardware_filtering: bool = False):
        """
        Args:
            width, height, config, light: as the reference (balanced preset / ``Light.default()``).
            device: CUDA device ordinal of this renderer's context.
            texel_format: ``"f32"`` ({s,nx,ny,nz} binary32, parity configs) or ``"f16"``.
            strict: reference-faithful arithmetic (slow; used to pin the kernel to the oracle).
            empty_space_skipping: exact macrocell skipping (pixels unchanged).
            hardware_filtering: sample through the texture unit (3-D CUDA array, hardware trilinear
                filter with 8-bit weights) instead of binary32 software trilinear.  Within the stated
                tolerance of the oracle but not bit-comparable with it; off by default.
            honor_config_termination: NON-PARITY opt-in -- stop rays at
                ``config.opacity_threshold`` (or never, if ``early_ray_termination`` is off)
                instead of the shader's literal 0.99.
        """
        self.width = width
        self.height = height
        self.volume: Optional[Volume] = None
        self.camera: Optional[Camera] = None

        if config is None:
            self.config = RenderConfig.balanced()
        else:
            if not _is_a(config, RenderConfig, "RenderConfig"):
                raise TypeError(f"Expected RenderConfig instance, got {type(config)}")
            self.config = config
        if light is None:
            self.light = Light.default()
        else:
            if not _is_a(light, Light, "Light"):
                raise TypeError(f"Expected Light instance, got {type(light)}")
            self.light = light

        if texel_format not in ("f32", "f16"):
            raise ValueError("texel_format must be 'f32' or 'f16'")
        self.device = int(device)
        self._texel_format = _cabi.TEXEL_F16X4 if texel_format == "f16" else _cabi.TEXEL_F32X4
        self._strict = bool(strict)
        self._ess = bool(empty_space_skipping)
        self._honor_termination = bool(honor_config_termination)
        self._hwtex = bool(hardware_filtering)
        self._lib = _cabi.lib()
        # the resource layer, driven through ModernGLManager's method names (renderer.py:89-105)
        self.gl_manager = CudaManager(width, height, device=self.device, texel_format=self._texel_format,
                                      flags=self._flags(), termination_alpha=self._termination_alpha())
        self._ctx = self.gl_manager.ctx
        self.gl_manager.load_shaders()
        self._update_render_config()
        self.gl_manager.set_uniform_vector("volume_min_bounds", (-0.5, -0.5, -0.5))
        self.gl_manager.set_uniform_vector("volume_max_bounds", (0.5, 0.5, 0.5))
        self._update_light()

    # -- resource life cycle -------------------------------------------------------------
    def close(self) -> None:
        manager = getattr(self, "gl_manager", None)
        if manager is not None and getattr(manager, "ctx", None):
            manager.cleanup()
        self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- reference API (same calls into the manager, in the same order, as renderer.py:107-316) ----------------
    def load_volume(self, volume: Volume) -> None:
        if not _is_a(volume, Volume, "Volume"):
            raise TypeError(
                f"Expected Volume instance, got {type(volume)}. "
                "Create a Volume instance: from pyvr.volume import Volume; "
                "volume = Volume(data=your_array)")
        self.volume = volume
        texture_unit = self.gl_manager.create_volume_texture(volume.data)
        self.gl_manager.set_uniform_int("volume_texture", texture_unit)
        self.gl_manager.set_uniform_vector("volume_min_bounds", tuple(volume.min_bounds))
        self.gl_manager.set_uniform_vector("volume_max_bounds", tuple(volume.max_bounds))
        if volume.has_normals:
            normal_unit = self.gl_manager.create_normal_texture(volume.normals)
            self.gl_manager.set_uniform_int("normal_volume", normal_unit)

    def set_camera(self, camera: Camera) -> None:
        if not _is_a(camera, Camera, "Camera"):
            raise TypeError(f"Expected Camera instance, got {type(camera)}")
        self.camera = camera
        aspect = self.width / self.height
        view_matrix = camera.get_view_matrix()
        projection_matrix = camera.get_projection_matrix(aspect)
        position, _ = camera.get_camera_vectors()
        self.gl_manager.set_uniform_matrix("view_matrix", view_matrix)
        self.gl_manager.set_uniform_matrix("projection_matrix", projection_matrix)
        self.gl_manager.set_uniform_vector("camera_pos", tuple(position))

    def set_light(self, light: Light) -> None:
        if not _is_a(light, Light, "Light"):
            raise TypeError(f"Expected Light instance, got {type(light)}")
        self.light = light
        self._update_light()

    def set_transfer_functions(self, color_transfer_function, opacity_transfer_function,
                               size: Optional[int] = None) -> None:
        rgba_tex_unit = self.gl_manager.create_rgba_transfer_function_texture(
            color_transfer_function, opacity_transfer_function, size)
        self.gl_manager.set_uniform_int("transfer_function_lut", rgba_tex_unit)

    def render(self) -> bytes:
        """Raw RGBA8 pixels, ``width*height*4`` bytes, bottom row first."""
        self.gl_manager.clear_framebuffer(0.0, 0.0, 0.0, 0.0)
        self.gl_manager.setup_blending()
        self.gl_manager.render_quad()
        return self.gl_manager.read_pixels()

    def render_to_pil(self, data=None):
        from PIL import Image

        if data is None:
            data = self.render()
        image = Image.frombytes("RGBA", (self.width, self.height), data)
        return image.transpose(Image.FLIP_TOP_BOTTOM)

    def set_config(self, config) -> None:
        if not _is_a(config, RenderConfig, "RenderConfig"):
            raise TypeError(f"Expected RenderConfig instance, got {type(config)}")
        self.config = config
        self._update_render_config()

    def get_config(self):
        return self.config

    def get_light(self):
        return self.light

    def get_volume(self) -> Optional[Volume]:
        return self.volume

    def get_camera(self) -> Optional[Camera]:
        return self.camera

    def _update_render_config(self):
        self.gl_manager.set_uniform_float("step_size", self.config.step_size)
        self.gl_manager.set_uniform_int("max_steps", self.config.max_steps)
        self.gl_manager.set_uniform_float("reference_step_size", self.config.reference_step_size)
        if hasattr(self.gl_manager, "set_march_options"):     # non-parity opt-in follows the config (see __init__)
            self.gl_manager.set_march_options(termination_alpha=self._termination_alpha())

    def _update_light(self):
        self.gl_manager.set_uniform_float("ambient_light", self.light.ambient_intensity)
        self.gl_manager.set_uniform_float("diffuse_light", self.light.diffuse_intensity)
        self.gl_manager.set_uniform_vector("light_position", tuple(self.light.position))
        self.gl_manager.set_uniform_vector("light_target", tuple(self.light.target))

    # -- additive API ----------------------------------------------------------------------
    def set_lut(self, rgba: np.ndarray) -> None:
        """Upload a ready ``(size, 4) float32`` RGBA table."""
        self.gl_manager.set_uniform_int("transfer_function_lut", self.gl_manager.create_lut_texture(rgba))

    def render_array(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """``render()`` into a ``(height, width, 4) uint8`` array (pinned or not); no ``bytes`` copy."""
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        if (not isinstance(out, np.ndarray) or out.dtype != np.uint8 or not out.flags.c_contiguous
                or not out.flags.writeable or out.nbytes < self.height * self.width * 4):
            raise ValueError("out must be a writeable C-contiguous uint8 array of at least height*width*4 bytes")
        self._sync()
        _cabi.check(self._lib.pyvr_cuda_render(self._ctx, out.ctypes.data, 0))
        return out

    def make_views(self, cameras: Iterable) -> ctypes.Array:
        aspect = self.width / self.height
        cams = list(cameras)
        views = (_cabi.View * len(cams))()
        for i, cam in enumerate(cams):
            views[i] = _cabi.view_from_camera(cam, aspect)
        return views

    def render_batch(self, cameras=None, *, views=None, out: Optional[np.ndarray] = None,
                     device_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        """Render many views in one call (the turntable path).

        ``out``: ``(n, height, width, 4) uint8`` host array (pinned for full PCIe rate); allocated if
        omitted.  ``device_ptr``: raw device pointer to receive the frames instead (nothing returned).
        Frame k's device->host copy overlaps the march of frame k+1.
        """
        if views is None:
            if cameras is None:
                raise ValueError("render_batch needs `cameras` or `views`")
            views = self.make_views(cameras)
        self._sync()
        n = len(views)
        if device_ptr is not None:
            _cabi.check(self._lib.pyvr_cuda_render_batch(self._ctx, views, n, ctypes.c_void_p(device_ptr), 1))
            return None
        if out is None:
            out = np.empty((n, self.height, self.width, 4), dtype=np.uint8)
        if (not isinstance(out, np.ndarray) or out.dtype != np.uint8 or not out.flags.c_contiguous
                or not out.flags.writeable or out.nbytes < n * self.height * self.width * 4):
            raise ValueError("out must be a writeable C-contiguous uint8 buffer of n*height*width*4 bytes")
        _cabi.check(self._lib.pyvr_cuda_render_batch(self._ctx, views, n, out.ctypes.data, 0))
        return out

    def load_brick(self, data: np.ndarray, normals: Optional[np.ndarray], global_shape, origin, own_lo, own_hi,
                   min_bounds, max_bounds) -> None:
        """Sort-last: load the sub-block ``data = whole[origin : origin + data.shape]`` (``data[ix,iy,iz]``,
        i.e. numpy axis k = world axis k) of a volume of ``global_shape`` voxels with bounds
        ``min_bounds..max_bounds``; this renderer then produces only the samples whose voxel coordinate
        lies in ``[own_lo, own_hi)`` (see ``pyvr_cuda_upload_brick``; ``pyvr_b200.multi_gpu.split_bricks``
        computes the triples, ghost layer included)."""
        data = np.ascontiguousarray(data, dtype=np.float32)
        if data.ndim != 3:
            raise ValueError("Volume data must be 3D")
        if normals is not None:
            normals = np.ascontiguousarray(normals, dtype=np.float32)
            if normals.shape != data.shape + (3,):
                raise ValueError("Normal volume must have 3 channels (last dimension).")
        i3 = lambda v: (ctypes.c_int * 3)(*[int(x) for x in v])
        bmin, bmax = _cabi.vec3(min_bounds), _cabi.vec3(max_bounds)
        _cabi.check(self._lib.pyvr_cuda_upload_brick(
            self._ctx, data.ctypes.data, normals.ctypes.data if normals is not None else None,
            i3(data.shape), i3(global_shape), i3(origin), i3(own_lo), i3(own_hi),
            _cabi.f32_ptr(bmin), _cabi.f32_ptr(bmax), self._texel_format, 0))
        self.gl_manager.adopt_device_volume()

    def generate_volume(self, size: int, shape: str = "double_sphere", min_bounds=(-0.5, -0.5, -0.5),
                        max_bounds=(0.5, 0.5, 0.5), brick=None) -> float:
        """Device-side equivalent of ``create_sample_volume(size, shape)`` + ``compute_normal_volume`` +
        ``load_volume`` (nothing crosses PCIe).  ``brick``: a ``pyvr_b200.multi_gpu.Brick`` to generate only
        that sub-block for sort-last rendering.  Returns the device time in milliseconds.  ``self.volume``
        stays ``None``: the data exists only in packed form on the device."""
        if shape not in _cabi.SHAPES:
            raise ValueError(f"Unknown shape: {shape}. Device-side shapes: {', '.join(_cabi.SHAPES)}")
        i3 = lambda v: (ctypes.c_int * 3)(*[int(x) for x in v])
        bmin, bmax = _cabi.vec3(min_bounds), _cabi.vec3(max_bounds)
        ms = ctypes.c_float(0.0)
        args = (i3(brick.dims), i3(brick.origin), i3(brick.own_lo), i3(brick.own_hi)) if brick is not None else (None,) * 4
        _cabi.check(self._lib.pyvr_cuda_generate_volume(
            self._ctx, _cabi.SHAPES[shape], int(size), *args, _cabi.f32_ptr(bmin), _cabi.f32_ptr(bmax),
            self._texel_format, ctypes.byref(ms)))
        self.volume = None
        self.gl_manager.adopt_device_volume()
        return ms.value

    def read_texels(self, shape) -> tuple:
        """Test aid: the stored block unpacked to ``(scalar (shape), normals (shape + (3,)))`` float32."""
        self._sync()
        scalar = np.empty(tuple(shape), dtype=np.float32)
        normals = np.empty(tuple(shape) + (3,), dtype=np.float32)
        _cabi.check(self._lib.pyvr_cuda_read_texels(self._ctx, scalar.ctypes.data, normals.ctypes.data))
        return scalar, normals

    def set_pixel_shard(self, rank: int, count: int, in_place: bool = False, group_shift: Optional[int] = None) -> None:
        """Image-space sharding: march only the tile groups of ``rank`` (of ``count``).  By default the rest of
        the frame is cleared, so the frames of all ranks add up to the full frame; ``in_place=True`` leaves the
        other pixels untouched (all ranks write one shared frame, ``multi_gpu.TileSession``).  ``group_shift``: tile
        groups of ``2^s x 2^s`` CTA tiles of 16x8 pixels (default 1)."""
        if group_shift is not None:
            _cabi.check(self._lib.pyvr_cuda_set_option(self._ctx, b"shard_shift", int(group_shift)))
        _cabi.check(self._lib.pyvr_cuda_set_option(self._ctx, b"shard_in_place", int(bool(in_place))))
        _cabi.check(self._lib.pyvr_cuda_set_pixel_shard(self._ctx, int(rank), int(count)))

    def render_to_device(self, device_ptr: int) -> None:
        """``render()`` into a device buffer of ``width*height*4`` bytes."""
        self._sync()
        _cabi.check(self._lib.pyvr_cuda_render(self._ctx, ctypes.c_void_p(device_ptr), 1))

    def render_accum_to_device(self, device_ptr: int) -> None:
        """Pre-blend fragment colours (``width*height`` float4) into a device buffer: a partial image."""
        self._sync()
        _cabi.check(self._lib.pyvr_cuda_render_accum(self._ctx, ctypes.c_void_p(device_ptr), 1))

    def render_tensor(self, cameras=None, out=None):
        """Zero-copy device output: the current view -- or, with ``cameras``, one
        frame per camera -- as a ``torch.uint8`` tensor ``([n,] height, width, 4)`` on this renderer's device.
        Nothing crosses PCIe except the view parameters; row 0 is the bottom row, as in ``render()``."""
        import torch

        n = None if cameras is None else len(cameras)
        shape = (self.height, self.width, 4) if n is None else (n, self.height, self.width, 4)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=f"cuda:{self.device}")
        if tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device.index != self.device:
            raise ValueError(f"out must be a contiguous uint8 CUDA tensor of shape {shape} on device {self.device}")
        if n is None:
            self._sync()
            _cabi.check(self._lib.pyvr_cuda_render(self._ctx, ctypes.c_void_p(out.data_ptr()), 1))
        else:
            self.render_batch(cameras, device_ptr=out.data_ptr())
        return out

    def render_accum_relay(self, in_ptr: Optional[int], out_ptr: int) -> None:
        """Sort-last relay: continue the device image ``in_ptr`` (fragment colours of the bricks in front;
        ``None`` for the first brick) through this renderer's brick into ``out_ptr`` (may be the same)."""
        self._sync()
        _cabi.check(self._lib.pyvr_cuda_render_accum_relay(
            self._ctx, ctypes.c_void_p(in_ptr) if in_ptr else None, ctypes.c_void_p(out_ptr)))

    def render_accum(self) -> np.ndarray:
        """Fragment colour before blending: ``(height, width, 4) float32`` = (acc_rgb, acc_a)."""
        self._sync()
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        _cabi.check(self._lib.pyvr_cuda_render_accum(self._ctx, out.ctypes.data, 0))
        return out

    def set_stream(self, cuda_stream: int) -> None:
        """Run this renderer's work on an existing ``cudaStream_t`` (e.g. torch's current stream).  ``0`` selects the
        renderer's own non-blocking stream; pass ``1`` (``cudaStreamLegacy``) for the legacy default stream."""
        _cabi.check(self._lib.pyvr_cuda_set_stream(self._ctx, ctypes.c_void_p(cuda_stream)))

    @property
    def texel_layout(self) -> str:
        """Layout of the packed texels of the loaded volume and how the march walks it (automatic choices resolved)."""
        def get(key):
            v = ctypes.c_int(0)
            _cabi.check(self._lib.pyvr_cuda_get_option(self._ctx, key, ctypes.byref(v)))
            return v.value
        base = "2x2x2-texel bricks" if get(b"brick8") else "rows of z-pair entries" if get(b"pair") else "rows"
        return base + (", several samples in flight per ray" if get(b"two_samples") else "")

    def set_async_device_output(self, enabled: bool) -> None:
        """``True``: renders into DEVICE buffers (``render_to_device``, ``render_accum_to_device``, ``render_batch(device_ptr=...)``,
        relay) return as soon as the work is enqueued -- the pixels are valid in the order of the renderer's stream,
        and ``stats`` waits for the counters.  Default ``False``: every render call ends with a host synchronisation.
        The multi-GPU sessions switch it on for the renderer they are bound to."""
        _cabi.check(self._lib.pyvr_cuda_set_option(self._ctx, b"async_device_output", int(bool(enabled))))

    @property
    def stats(self) -> dict:
        """Work counters and device time of the last render call (``pyvr_stats``)."""
        s = _cabi.Stats()
        _cabi.check(self._lib.pyvr_cuda_get_stats(self._ctx, ctypes.byref(s)))
        return s.as_dict()

    # -- internals -------------------------------------------------------------------------
    def _termination_alpha(self) -> float:
        if not self._honor_termination:
            return REFERENCE_TERMINATION_ALPHA
        if not self.config.early_ray_termination:
            return 2.0  # never reached
        return float(self.config.opacity_threshold)

    def _flags(self) -> int:
        flags = 0
        if self._strict:
            flags |= _cabi.FLAG_STRICT
        elif self._ess:
            flags |= _cabi.FLAG_ESS
        if self._hwtex and not self._strict:
            flags |= _cabi.FLAG_HWTEX
        return flags

    def _sync(self) -> None:
        """Additive entry points call the C ABI directly: first bring the device in line with what the reference
        API recorded in the manager (volume, camera, uniforms)."""
        self.gl_manager.flush()


# Same alias the reference exports (renderer.py:320).
VolumeRenderer = CudaVolumeRenderer
