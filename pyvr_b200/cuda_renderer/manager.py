"""``CudaManager`` -- ``ModernGLManager``'s interface (pyvr/moderngl_renderer/manager.py:14-256) over the C ABI.

The reference's renderer talks to its GPU resource layer through a dozen calls -- ``create_volume_texture``,
``create_normal_texture``, ``create_rgba_transfer_function_texture``, ``set_uniform_{matrix,vector,float,int}``,
``clear_framebuffer``, ``setup_blending``, ``render_quad``, ``read_pixels``, ``cleanup`` -- and its tests replace
exactly those with mocks (tests/test_moderngl_renderer/test_volume_renderer.py).  ``VolumeRenderer`` here drives
its ``gl_manager`` through the same calls with the same arguments, so that code written against the reference
class -- including those tests and the matplotlib front end -- runs unchanged on the CUDA backend.

There are no textures or uniforms underneath: the calls record state, and ``render_quad`` flushes what changed
to ``libpyvr_cuda.so`` (one packed texel upload per volume, camera, parameters) before it launches the march.
"Texture units" are kept as the reference keeps them (a counter that only grows, manager.py:232-236) because they
carry one piece of behaviour: a Volume without normals never binds ``normal_volume``, the sampler stays on unit
0 = the scalar texture, and the shader shades with ``normal = (density, 0, 0)`` (renderer.py:143-146).
"""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _cabi


class CudaManager:
    def __init__(self, width=512, height=512, *, device: int = 0, texel_format: int = _cabi.TEXEL_F32X4, flags: int = _cabi.FLAG_ESS,
                 termination_alpha: float = 0.99):
        self.width, self.height, self.device = int(width), int(height), int(device)
        self._texel_format, self._flags, self._termination_alpha = texel_format, int(flags), float(termination_alpha)
        self._lib = _cabi.lib()
        self.ctx = ctypes.c_void_p()
        _cabi.check(self._lib.pyvr_cuda_create(self.device, self.width, self.height, ctypes.byref(self.ctx)))
        self.program = None                    # set by load_shaders, as in the reference
        self._next_texture_unit = 0
        self._textures = {}                    # unit -> ("scalar" | "normal" | "lut", array or None)
        self._uniforms = {"volume_texture": 0, "normal_volume": 0, "transfer_function_lut": 0,
                          "volume_min_bounds": (-0.5, -0.5, -0.5), "volume_max_bounds": (0.5, 0.5, 0.5)}
        self._dirty = {"volume": False, "camera": False, "params": True}
        self._uploaded = None                  # (scalar unit, normal unit or None, bounds) of the texels on the device
        self._blend = False
        self._staging = None

    # ---- shaders: nothing to compile, the march kernels are in the library ---------------------------------
    def load_shaders(self, vertex_shader_path=None, fragment_shader_path=None):
        self.program = "libpyvr_cuda.so:march_kernel"

    # ---- textures ----------------------------------------------------------------------------------------------
    def _get_next_texture_unit(self):
        unit = self._next_texture_unit
        self._next_texture_unit += 1
        return unit

    def create_volume_texture(self, volume_data):
        if len(volume_data.shape) != 3:
            raise ValueError("Volume data must be 3D")
        if volume_data.dtype != np.float32:
            volume_data = volume_data.astype(np.float32)
        unit = self._get_next_texture_unit()
        self._textures[unit] = ("scalar", np.ascontiguousarray(volume_data))
        self._dirty["volume"] = True
        return unit

    def create_normal_texture(self, normal_data):
        if normal_data.shape[-1] != 3:
            raise ValueError("Normal volume must have 3 channels (last dimension).")
        unit = self._get_next_texture_unit()
        self._textures[unit] = ("normal", np.ascontiguousarray(normal_data, dtype=np.float32))
        self._dirty["volume"] = True
        return unit

    def create_rgba_transfer_function_texture(self, color_transfer_function, opacity_transfer_function,
                                              size: Optional[int] = None) -> int:
        from ..transferfunctions import build_rgba_lut

        lut = np.ascontiguousarray(build_rgba_lut(color_transfer_function, opacity_transfer_function, size), dtype=np.float32)
        return self.create_lut_texture(lut)

    def create_lut_texture(self, rgba: np.ndarray) -> int:
        """A ready ``(size, 4) float32`` table (no reference counterpart; what the method above builds)."""
        lut = np.ascontiguousarray(rgba, dtype=np.float32)
        if lut.ndim != 2 or lut.shape[1] != 4:
            raise ValueError("LUT must have shape (size, 4)")
        _cabi.check(self._lib.pyvr_cuda_set_lut(self.ctx, lut.ctypes.data, lut.shape[0]))
        unit = self._get_next_texture_unit()
        self._textures[unit] = ("lut", None)
        return unit

    # ---- uniforms ------------------------------------------------------------------------------------------------
    def _set(self, name, value):
        if self.program is None:
            raise RuntimeError("Shader program not loaded")
        self._uniforms[name] = value
        if name in ("view_matrix", "projection_matrix", "camera_pos"):
            self._dirty["camera"] = True
        elif name in ("volume_texture", "normal_volume", "volume_min_bounds", "volume_max_bounds"):
            self._dirty["volume"] = True
        else:
            self._dirty["params"] = True

    def set_uniform_matrix(self, name, matrix):
        self._set(name, np.ascontiguousarray(matrix, dtype=np.float32).reshape(16).copy())   # .write(matrix.tobytes())

    def set_uniform_vector(self, name, vector):
        self._set(name, tuple(float(v) for v in vector))

    def set_uniform_float(self, name, value):
        self._set(name, float(value))

    def set_uniform_int(self, name, value):
        self._set(name, int(value))

    # ---- draw ----------------------------------------------------------------------------------------------------
    def clear_framebuffer(self, r=0.0, g=0.0, b=0.0, a=0.0):
        if (r, g, b, a) != (0.0, 0.0, 0.0, 0.0):
            raise NotImplementedError("the march blends onto the reference's (0,0,0,0) clear (renderer.py:216)")

    def setup_blending(self):
        self._blend = True

    def set_march_options(self, flags: Optional[int] = None, termination_alpha: Optional[float] = None):
        if flags is not None:
            self._flags = int(flags)
        if termination_alpha is not None:
            self._termination_alpha = float(termination_alpha)
        self._dirty["params"] = True

    def adopt_device_volume(self):
        """The texels on the device were produced without textures (a sort-last brick, a device-side synthetic
        volume): forget the recorded scalar / normal textures so that the next flush does not replace them."""
        unit = self._get_next_texture_unit()
        self._textures[unit] = ("device", None)
        self._uniforms["volume_texture"] = unit
        self._uniforms["normal_volume"] = unit
        self._uploaded = None
        self._dirty["volume"] = False

    def flush(self):
        """Bring the device in line with the recorded textures and uniforms."""
        u = self._uniforms
        if self._dirty["volume"]:
            kind, scalar = self._textures.get(u["volume_texture"], (None, None))
            if kind == "scalar":
                n_kind, normals = self._textures.get(u["normal_volume"], (None, None))
                if n_kind != "normal":
                    normals = None            # sampler left on the scalar texture: normal = (density, 0, 0)
                elif normals.shape[:3] != scalar.shape:
                    raise RuntimeError("normal_volume is bound to a normal texture of another size (a stale texture of a "
                                       "previously loaded Volume); load a Volume with its own normals")
                key = (u["volume_texture"], u["normal_volume"] if normals is not None else None,
                       u["volume_min_bounds"], u["volume_max_bounds"])
                if key != self._uploaded:
                    bmin, bmax = _cabi.vec3(u["volume_min_bounds"]), _cabi.vec3(u["volume_max_bounds"])
                    _cabi.check(self._lib.pyvr_cuda_upload_volume(
                        self.ctx, scalar.ctypes.data, normals.ctypes.data if normals is not None else None,
                        scalar.shape[0], scalar.shape[1], scalar.shape[2], _cabi.f32_ptr(bmin), _cabi.f32_ptr(bmax),
                        self._texel_format, 0))
                    self._uploaded = key
                    # the device holds the texels now: drop the host copies of every texture but the bound ones
                    keep = {u["volume_texture"], u["normal_volume"]}
                    self._textures = {k: (v if k in keep or v[0] == "lut" else (v[0], None)) for k, v in self._textures.items()}
            self._dirty["volume"] = False
        if self._dirty["camera"] and all(k in u for k in ("view_matrix", "projection_matrix", "camera_pos")):
            pos = _cabi.vec3(u["camera_pos"])
            _cabi.check(self._lib.pyvr_cuda_set_camera(self.ctx, _cabi.f32_ptr(u["view_matrix"]),
                                                       _cabi.f32_ptr(u["projection_matrix"]), _cabi.f32_ptr(pos)))
            self._dirty["camera"] = False
        if self._dirty["params"]:
            p = _cabi.Params()
            p.step_size = u.get("step_size", 0.01)
            p.max_steps = u.get("max_steps", 500)
            p.reference_step_size = u.get("reference_step_size", 0.01)
            p.ambient, p.diffuse = u.get("ambient_light", 0.2), u.get("diffuse_light", 0.8)
            p.light_position[:] = u.get("light_position", (1.0, 1.0, 1.0))
            p.light_target[:] = u.get("light_target", (0.0, 0.0, 0.0))
            p.termination_alpha = self._termination_alpha
            p.flags = self._flags
            _cabi.check(self._lib.pyvr_cuda_set_params(self.ctx, ctypes.byref(p)))
            self._dirty["params"] = False

    def render_quad(self):
        """The draw call: march every pixel into the RGBA8 target (and bring it to the host, where the reference's
        ``fbo.read`` would)."""
        if self.program is None:
            raise RuntimeError("Vertex array object not created")
        self.flush()
        if self._staging is None:      # page-locked read-back target: full PCIe rate, one copy into the bytes object
            self._staging = _cabi.PinnedBuffer(self.width * self.height * 4)
        _cabi.check(self._lib.pyvr_cuda_render(self.ctx, self._staging.array.ctypes.data, 0))

    def read_pixels(self):
        """``fbo.read(components=4)``: ``width*height*4`` bytes, bottom row first."""
        if self._staging is None:
            return bytes(self.width * self.height * 4)
        return self._staging.array.tobytes()

    def cleanup(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        if ctx:
            self._lib.pyvr_cuda_destroy(ctx)
        staging, self._staging = getattr(self, "_staging", None), None
        if staging is not None:
            staging.close()
        self._textures = {}
