"""ctypes binding of ``libpyvr_cuda.so`` (the C ABI declared in ``include/pyvr_cuda.h``).

The library is built in-tree by :func:`pyvr_b200._build.build_library` (``nvcc`` for sm_100a).  There
is no CPU fallback: if the library cannot be loaded, or no CUDA device is visible, every compute
entry point raises ``RuntimeError``.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libpyvr_cuda.so")

ABI_VERSION = 2
IPC_HANDLE_BYTES = 64
TEXEL_F32X4, TEXEL_F16X4 = 0, 1
FLAG_STRICT, FLAG_ESS, FLAG_NO_BLEND, FLAG_HWTEX = 0x1, 0x2, 0x4, 0x8
SHAPES = {"sphere": 0, "torus": 1, "double_sphere": 2}

# name -> (restype, argtypes); every symbol include/pyvr_cuda.h declares
_c = ctypes
_vp, _i, _fp, _ip = _c.c_void_p, _c.c_int, _c.POINTER(_c.c_float), _c.POINTER(_c.c_int)


class View(ctypes.Structure):
    """``pyvr_view``: ray origin, closed-form basis and the binary32 inverse matrices of one view."""
    _fields_ = [("origin", _c.c_float * 3), ("u", _c.c_float * 3), ("v", _c.c_float * 3), ("w", _c.c_float * 3),
                ("inv_proj", _c.c_float * 16), ("inv_view", _c.c_float * 16), ("has_matrices", _c.c_int32)]


class Params(ctypes.Structure):
    """``pyvr_params``: the shader's non-camera uniforms."""
    _fields_ = [
        ("step_size", _c.c_float), ("max_steps", _c.c_int32), ("reference_step_size", _c.c_float),
        ("ambient", _c.c_float), ("diffuse", _c.c_float),
        ("light_position", _c.c_float * 3), ("light_target", _c.c_float * 3),
        ("termination_alpha", _c.c_float), ("flags", _c.c_uint32),
    ]


class Stats(ctypes.Structure):
    """``pyvr_stats``."""
    _fields_ = [
        ("samples", _c.c_uint64), ("samples_fetched", _c.c_uint64), ("rays_hit", _c.c_uint64),
        ("rays_terminated", _c.c_uint64), ("kernel_ms", _c.c_float),
        ("kernel_launches", _c.c_uint32), ("views", _c.c_uint32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


SYMBOLS = {
    "pyvr_cuda_create": (_i, [_i, _i, _i, _c.POINTER(_vp)]),
    "pyvr_cuda_destroy": (_i, [_vp]),
    "pyvr_cuda_set_stream": (_i, [_vp, _vp]),
    "pyvr_cuda_upload_volume": (_i, [_vp, _vp, _vp, _i, _i, _i, _fp, _fp, _i, _i]),
    "pyvr_cuda_upload_brick": (_i, [_vp, _vp, _vp, _ip, _ip, _ip, _ip, _ip, _fp, _fp, _i, _i]),
    "pyvr_cuda_generate_volume": (_i, [_vp, _i, _i, _ip, _ip, _ip, _ip, _fp, _fp, _i, _fp]),
    "pyvr_cuda_read_texels": (_i, [_vp, _vp, _vp]),
    "pyvr_cuda_set_pixel_shard": (_i, [_vp, _i, _i]),
    "pyvr_cuda_set_lut": (_i, [_vp, _vp, _i]),
    "pyvr_cuda_set_camera": (_i, [_vp, _fp, _fp, _fp]),
    "pyvr_cuda_view_from_matrices": (_i, [_fp, _fp, _fp, _c.POINTER(View)]),
    "pyvr_cuda_set_view": (_i, [_vp, _c.POINTER(View)]),
    "pyvr_cuda_set_params": (_i, [_vp, _c.POINTER(Params)]),
    "pyvr_cuda_render": (_i, [_vp, _vp, _i]),
    "pyvr_cuda_render_batch": (_i, [_vp, _vp, _i, _vp, _i]),
    "pyvr_cuda_render_accum": (_i, [_vp, _vp, _i]),
    "pyvr_cuda_render_accum_relay": (_i, [_vp, _vp, _vp]),
    "pyvr_cuda_get_stats": (_i, [_vp, _c.POINTER(Stats)]),
    "pyvr_cuda_compute_normals": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _fp]),
    "pyvr_cuda_composite_over": (_i, [_i, _vp, _vp, _vp, _c.c_size_t, _c.c_float, _vp]),
    "pyvr_cuda_finalize_rgba8": (_i, [_i, _vp, _vp, _c.c_size_t, _c.c_uint32, _vp]),
    "pyvr_cuda_composite_finalize": (_i, [_i, _vp, _vp, _vp, _vp, _c.c_size_t, _c.c_float, _c.c_uint32, _vp]),
    "pyvr_cuda_flag_signal": (_i, [_i, _vp, _c.c_uint32, _vp]),
    "pyvr_cuda_flag_wait": (_i, [_i, _vp, _i, _c.c_uint32, _vp]),
    "pyvr_cuda_flag_signal_many": (_i, [_i, _vp, _i, _c.c_uint32, _vp]),
    "pyvr_cuda_binary_swap": (_i, [_i, _vp, _i, _c.c_uint32, _c.c_float, _c.c_uint32, _vp]),
    "pyvr_cuda_device_alloc": (_i, [_i, _c.c_size_t, _c.POINTER(_vp)]),
    "pyvr_cuda_device_free": (_i, [_i, _vp]),
    "pyvr_cuda_ipc_export": (_i, [_i, _vp, _vp]),
    "pyvr_cuda_ipc_open": (_i, [_i, _vp, _c.POINTER(_vp)]),
    "pyvr_cuda_ipc_close": (_i, [_i, _vp]),
    "pyvr_cuda_memcpy": (_i, [_i, _vp, _vp, _c.c_size_t, _i, _vp]),
    "pyvr_cuda_stream_synchronize": (_i, [_i, _vp]),
    "pyvr_cuda_set_option": (_i, [_vp, _c.c_char_p, _i]),
    "pyvr_cuda_get_option": (_i, [_vp, _c.c_char_p, _c.POINTER(_i)]),
    "pyvr_cuda_measure_cache_bandwidth": (_i, [_i, _i, _c.POINTER(_c.c_double)]),
    "pyvr_cuda_host_alloc": (_i, [_c.c_size_t, _c.POINTER(_vp)]),
    "pyvr_cuda_host_free": (_i, [_vp]),
    "pyvr_cuda_device_count": (_i, [_c.POINTER(_i)]),
    "pyvr_cuda_abi_version": (_i, []),
    "pyvr_cuda_last_error": (_c.c_char_p, []),
}

_lib = None


def lib():
    """Load ``libpyvr_cuda.so`` (once) and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PYVR_CUDA_LIB") or LIB_PATH      # override: A/B builds of the same ABI
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  pyvr_b200 has no CPU fallback.")
    handle = ctypes.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = restype, argtypes
    if handle.pyvr_cuda_abi_version() != ABI_VERSION:
        raise RuntimeError("libpyvr_cuda.so ABI version mismatch; rebuild the library")
    _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().pyvr_cuda_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"pyvr_cuda error {rc}: {last_error()}")


def device_count() -> int:
    n = _c.c_int(0)
    check(lib().pyvr_cuda_device_count(_c.byref(n)))
    return n.value


def measure_cache_bandwidth(level: int, device: int = 0) -> float:
    """GB/s delivered to registers by coalesced 128-bit loads served from L1 (``level=1``) or L2 (``level=2``)."""
    gbs = _c.c_double(0.0)
    check(lib().pyvr_cuda_measure_cache_bandwidth(int(device), int(level), _c.byref(gbs)))
    return gbs.value


def f32_ptr(a: np.ndarray):
    return a.ctypes.data_as(_fp)


def vec3(values) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(3))
    return a


def view_from_matrices(view: np.ndarray, proj: np.ndarray, cam_pos) -> View:
    """``pyvr_view`` from the reference's three camera uniforms (matrices as ``.tobytes()`` order)."""
    v = np.ascontiguousarray(view, dtype=np.float32).reshape(16)
    p = np.ascontiguousarray(proj, dtype=np.float32).reshape(16)
    pos = vec3(cam_pos)
    out = View()
    check(lib().pyvr_cuda_view_from_matrices(f32_ptr(v), f32_ptr(p), f32_ptr(pos), _c.byref(out)))
    return out


def view_from_camera(camera, aspect: float) -> View:
    position, _ = camera.get_camera_vectors()
    return view_from_matrices(camera.get_view_matrix(), camera.get_projection_matrix(aspect), position)


NORMALS_DEVICE_BUFFERS, NORMALS_RELAXED = 1, 2     # pyvr_cuda_compute_normals flags


def compute_normals_host(volume: np.ndarray, device: int = 0, return_ms: bool = False, relaxed: bool = False):
    """Host array in, host array out, through ``pyvr_cuda_compute_normals``."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    out = np.empty(vol.shape + (3,), dtype=np.float32)
    ms = _c.c_float(0.0)
    check(lib().pyvr_cuda_compute_normals(device, vol.ctypes.data, out.ctypes.data,
                                          vol.shape[0], vol.shape[1], vol.shape[2],
                                          NORMALS_RELAXED if relaxed else 0, _c.byref(ms)))
    return (out, ms.value) if return_ms else out


class DeviceBuffer:
    """Plain ``cudaMalloc`` memory (IPC-exportable).  ``ptr`` is the raw device address."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = int(device), int(nbytes)
        p = _vp()
        check(lib().pyvr_cuda_device_alloc(self.device, max(self.nbytes, 1), _c.byref(p)))
        self.ptr = p.value

    def to_host(self, dtype=np.uint8, offset: int = 0, nbytes: Optional[int] = None) -> np.ndarray:
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(nbytes, dtype=np.uint8)
        check(lib().pyvr_cuda_memcpy(self.device, out.ctypes.data, _vp(self.ptr + offset), nbytes, 2, None))
        return out.view(dtype)

    def from_host(self, array: np.ndarray, offset: int = 0) -> None:
        a = np.ascontiguousarray(array)
        check(lib().pyvr_cuda_memcpy(self.device, _vp(self.ptr + offset), a.ctypes.data, a.nbytes, 1, None))

    def ipc_handle(self) -> bytes:
        h = (_c.c_uint8 * IPC_HANDLE_BYTES)()
        check(lib().pyvr_cuda_ipc_export(self.device, _vp(self.ptr), h))
        return bytes(h)

    def close(self) -> None:
        if getattr(self, "ptr", None):
            lib().pyvr_cuda_device_free(self.device, _vp(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBuffer:
    """A ``DeviceBuffer`` of another process on this node, mapped through CUDA IPC."""

    def __init__(self, handle: bytes, device: int = 0):
        self.device = int(device)
        h = (_c.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(handle)
        p = _vp()
        check(lib().pyvr_cuda_ipc_open(self.device, h, _c.byref(p)))
        self.ptr = p.value

    def close(self) -> None:
        if getattr(self, "ptr", None):
            lib().pyvr_cuda_ipc_close(self.device, _vp(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def composite_over(device: int, front_ptr: int, back_ptr: int, out_ptr: int, n_pixels: int,
                   termination_alpha: float = 0.99, stream: int = 0) -> None:
    """``out = front over back`` on device pointers to ``n_pixels`` float4 pre-blend fragment colours."""
    check(lib().pyvr_cuda_composite_over(device, _vp(front_ptr), _vp(back_ptr), _vp(out_ptr), n_pixels,
                                         termination_alpha, _vp(stream)))


def finalize_rgba8(device: int, accum_ptr: int, out_ptr: int, n_pixels: int, flags: int = 0, stream: int = 0) -> None:
    check(lib().pyvr_cuda_finalize_rgba8(device, _vp(accum_ptr), _vp(out_ptr), n_pixels, flags, _vp(stream)))


def composite_finalize(device: int, front_ptr: int, back_ptr: int, accum_out_ptr: Optional[int], out_ptr: int, n_pixels: int,
                       termination_alpha: float = 0.99, flags: int = 0, stream: int = 0) -> None:
    """Last binary-swap round fused with finalize: ``out = rgba8(front over back)`` (``out`` may be peer memory)."""
    check(lib().pyvr_cuda_composite_finalize(device, _vp(front_ptr), _vp(back_ptr), _vp(accum_out_ptr) if accum_out_ptr else None,
                                             _vp(out_ptr), n_pixels, termination_alpha, flags, _vp(stream)))


def flag_signal(device: int, flag_ptr: int, value: int, stream: int = 0) -> None:
    check(lib().pyvr_cuda_flag_signal(device, _vp(flag_ptr), value & 0xFFFFFFFF, _vp(stream)))


def flag_signal_many(device: int, flag_ptrs, value: int, stream: int = 0) -> None:
    """The same release store to several flags (one per peer) in one launch."""
    arr = (_vp * len(flag_ptrs))(*[_vp(p) for p in flag_ptrs])
    check(lib().pyvr_cuda_flag_signal_many(device, arr, len(flag_ptrs), value & 0xFFFFFFFF, _vp(stream)))


class SwapRound(_c.Structure):
    """``pyvr_swap_round`` (include/pyvr_cuda.h): one round of a binary swap, raw device pointers (0 = NULL)."""
    _fields_ = [("front", _vp), ("back", _vp), ("out", _vp), ("out8", _vp), ("n_pixels", _c.c_uint64),
                ("signal_before", _vp), ("wait_flag", _vp), ("signal_done", _vp), ("signal_next", _vp)]


def binary_swap(device: int, rounds, value: int, termination_alpha: float = 0.99, flags: int = 0, stream: int = 0) -> None:
    """All rounds of a binary swap enqueued back to back by one C call (``pyvr_cuda_binary_swap``)."""
    arr = (SwapRound * len(rounds))(*rounds)
    check(lib().pyvr_cuda_binary_swap(device, arr, len(rounds), value & 0xFFFFFFFF, termination_alpha, flags, _vp(stream)))


def flag_wait(device: int, flags_ptr: int, n_flags: int, value: int, stream: int = 0) -> None:
    check(lib().pyvr_cuda_flag_wait(device, _vp(flags_ptr), n_flags, value & 0xFFFFFFFF, _vp(stream)))


def stream_synchronize(device: int, stream: int = 0) -> None:
    check(lib().pyvr_cuda_stream_synchronize(device, _vp(stream)))


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy ``uint8`` array (frame read-back target)."""

    def __init__(self, nbytes: int):
        self._ptr = _vp()
        check(lib().pyvr_cuda_host_alloc(nbytes, _c.byref(self._ptr)))
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((_c.c_uint8 * nbytes).from_address(self._ptr.value))

    def close(self) -> None:
        if self._ptr:
            self.array = None
            lib().pyvr_cuda_host_free(self._ptr)
            self._ptr = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
