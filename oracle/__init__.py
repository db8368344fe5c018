"""CPU parity oracle for the ray-march path.  TEST INFRASTRUCTURE ONLY.

ctypes front end of ``oracle/liboracle.so`` (built from ``oracle/pyvr_oracle.c`` by
``oracle/Makefile``).  Importers allowed: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg.  Nothing under ``pyvr_b200/`` may
import this package (``tests/test_no_oracle_in_product.py`` enforces it).

PARITY STATUS: pinned.  ``oracle.gl`` runs the reference's shader files verbatim on Mesa llvmpipe (see its
docstring) and ``tests/test_gl_reference.py`` holds this restatement to it (max |delta| = 1/255 on 19 scenes;
``profiles/r02_gl_pin.json``).  Normals, camera, LUT and sample-volume inputs are pinned by ``tests/golden/``.

The scene is assembled the way the reference renderer pushes its uniforms:
bounds ``renderer.py:136-141``, matrices/camera position ``renderer.py:164-172``,
step/max_steps/reference step ``renderer.py:303-309``, light ``renderer.py:311-316``,
textures ``manager.py:87-101,120-129,163-181``.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class _Scene(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("scalar", ctypes.c_void_p), ("normals", ctypes.c_void_p),
        ("tex_w", ctypes.c_int32), ("tex_h", ctypes.c_int32), ("tex_d", ctypes.c_int32),
        ("bmin", ctypes.c_float * 3), ("bmax", ctypes.c_float * 3),
        ("lut", ctypes.c_void_p), ("lut_size", ctypes.c_int32),
        ("view", ctypes.c_float * 16), ("proj", ctypes.c_float * 16),
        ("cam_pos", ctypes.c_float * 3),
        ("step_size", ctypes.c_float), ("max_steps", ctypes.c_int32), ("ref_step", ctypes.c_float),
        ("ambient", ctypes.c_float), ("diffuse", ctypes.c_float),
        ("light_pos", ctypes.c_float * 3), ("light_target", ctypes.c_float * 3),
    ]


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` if it is missing (or ``force``); returns its path."""
    src = os.path.join(_HERE, "pyvr_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       capture_output=True, text=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_render.restype = ctypes.c_int
        _lib.oracle_render.argtypes = [ctypes.POINTER(_Scene), ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_normals.restype = ctypes.c_int
        _lib.oracle_normals.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def render_scene(*, width: int, height: int, scalar: np.ndarray, normals: Optional[np.ndarray],
                 bmin, bmax, lut: np.ndarray, view: np.ndarray, proj: np.ndarray, cam_pos,
                 step_size: float, max_steps: int, ref_step: float, ambient: float, diffuse: float,
                 light_pos, light_target, rows: Tuple[int, int, int] = None,
                 want_accum: bool = False):
    """Low-level entry: raw arrays in, ``(rgba8 (H,W,4) uint8, accum or None, counters)`` out.

    ``counters`` = dict(samples, rays_hit, rays_terminated).  ``rows=(y0, y1, stride)`` restricts
    the rendered rows (used for the bounded CPU baseline); other rows stay zero.
    """
    scalar = _f32(scalar)
    if scalar.ndim != 3:
        raise ValueError("Volume data must be 3D")
    if normals is not None:
        normals = _f32(normals)
        if normals.shape != scalar.shape + (3,):
            raise ValueError("Normal volume must have 3 channels (last dimension).")
    lut = _f32(lut)
    sc = _Scene()
    sc.width, sc.height = int(width), int(height)
    sc.scalar = scalar.ctypes.data
    sc.normals = normals.ctypes.data if normals is not None else None
    # moderngl texture3d(volume_data.shape, ...): (width, height, depth) = shape, manager.py:95-97
    sc.tex_w, sc.tex_h, sc.tex_d = (int(s) for s in scalar.shape)
    sc.bmin[:] = [float(np.float32(v)) for v in bmin]
    sc.bmax[:] = [float(np.float32(v)) for v in bmax]
    sc.lut = lut.ctypes.data
    sc.lut_size = int(lut.shape[0])
    sc.view[:] = _f32(view).ravel().tolist()   # matrix.tobytes(), manager.py:192
    sc.proj[:] = _f32(proj).ravel().tolist()
    sc.cam_pos[:] = [float(np.float32(v)) for v in cam_pos]
    sc.step_size, sc.max_steps, sc.ref_step = float(step_size), int(max_steps), float(ref_step)
    sc.ambient, sc.diffuse = float(ambient), float(diffuse)
    sc.light_pos[:] = [float(np.float32(v)) for v in light_pos]
    sc.light_target[:] = [float(np.float32(v)) for v in light_target]

    out = np.zeros((height, width, 4), dtype=np.uint8)
    accum = np.zeros((height, width, 4), dtype=np.float32) if want_accum else None
    counters = np.zeros(3, dtype=np.uint64)
    y0, y1, ys = rows if rows is not None else (0, height, 1)
    rc = lib().oracle_render(ctypes.byref(sc), out.ctypes.data,
                             accum.ctypes.data if accum is not None else None,
                             int(y0), int(y1), int(ys), counters.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed ({rc})")
    stats = {"samples": int(counters[0]), "rays_hit": int(counters[1]),
             "rays_terminated": int(counters[2])}
    return out, accum, stats


def render(volume, camera, light, config, lut: np.ndarray, width: int, height: int,
           rows=None, want_accum: bool = False):
    """Render host objects (``Volume``, ``Camera``, ``Light``, ``RenderConfig``) + RGBA LUT.

    Accepts the repo's mirror classes or the reference's own (duck-typed).  Returns
    ``(rgba8, accum, counters)``; ``rgba8.tobytes()`` is what ``VolumeRenderer.render()`` returns
    (row 0 = bottom of the image).
    """
    position, _ = camera.get_camera_vectors()
    return render_scene(
        width=width, height=height, scalar=volume.data, normals=volume.normals,
        bmin=volume.min_bounds, bmax=volume.max_bounds, lut=lut,
        view=camera.get_view_matrix(), proj=camera.get_projection_matrix(width / height),
        cam_pos=position, step_size=config.step_size, max_steps=config.max_steps,
        ref_step=config.reference_step_size, ambient=light.ambient_intensity,
        diffuse=light.diffuse_intensity, light_pos=light.position, light_target=light.target,
        rows=rows, want_accum=want_accum)


def normals(volume: np.ndarray) -> np.ndarray:
    """``compute_normal_volume`` restated in C (binary32); ``(D,H,W) -> (D,H,W,3)``."""
    vol = _f32(volume)
    if vol.ndim != 3:
        raise ValueError("Volume data must be 3D")
    out = np.empty(vol.shape + (3,), dtype=np.float32)
    rc = lib().oracle_normals(vol.ctypes.data, out.ctypes.data, *[int(s) for s in vol.shape])
    if rc != 0:
        raise RuntimeError(f"oracle_normals failed ({rc})")
    return out
