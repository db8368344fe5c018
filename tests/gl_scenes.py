"""Scenes on which the CPU oracle is pinned to a REAL run of the reference's shader (Mesa llvmpipe through
oracle/gl).  Shared by tests/test_gl_reference.py and tests/golden/make_gl_golden.py, so the committed golden
frames and the live comparison are the same pictures.  Each entry: name -> (volume, camera, light, config,
lut, width, height)."""

import numpy as np

import oracle
from pyvr_b200 import (Camera, Light, RenderConfig, Volume, create_sample_volume)

from scenes import c1_scene, viridis_lut


def _pm1(data, normals):
    return Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32),
                  max_bounds=np.array([1, 1, 1], np.float32))


def scenes(small=False):
    """`small` = the subset stored as golden frames (tests/golden/gl_frames.npz)."""
    out = {}
    n = 64
    data = create_sample_volume(n, "double_sphere")
    normals = oracle.normals(data)
    vol, light, lut = c1_scene(n, normals=normals)
    # C1 (SURVEY.md section 8 d): balanced preset, directional light, viridis + linear(0, 0.3)
    for name, cam in (("c1_iso", Camera.isometric_view(distance=3.0)), ("c1_front", Camera.front_view(distance=3.0)),
                      ("c1_side", Camera.side_view(distance=3.0)), ("c1_top", Camera.top_view(distance=3.0)),
                      ("c1_rolled", Camera(azimuth=0.7, elevation=0.35, roll=0.5, distance=2.2)),
                      ("c1_inside", Camera(azimuth=0.3, elevation=0.2, roll=0.0, distance=0.3))):
        out[name] = (vol, cam, light, RenderConfig.balanced(), lut, 160, 128)
    # a Volume without normals: `normal_volume` stays on texture unit 0 = the scalar texture (renderer.py:143-146)
    vol_nn, _, _ = c1_scene(n, normals=None)
    out["no_normals"] = (vol_nn, Camera.isometric_view(distance=3.0), light, RenderConfig.balanced(), lut, 128, 128)
    # C3-like: bounds +-1, high_quality, turntable view, linear(0, 0.1); and the presets whose reach
    # step*max_steps = 2.0 is shorter than the chord (fast, ultra): the loop bound is max_steps, not t_far
    torus = create_sample_volume(48, "torus")
    vol3 = _pm1(torus, oracle.normals(torus))
    cam3 = Camera.from_spherical(target=np.zeros(3, np.float32), azimuth=2 * np.pi * 40 / 360, elevation=np.pi / 6,
                                 roll=0.0, distance=3.0)
    for preset in ("preview", "fast", "balanced", "high_quality", "ultra_quality"):
        out[f"pm1_{preset}"] = (vol3, cam3, Light.directional([1, -1, 0]), getattr(RenderConfig, preset)(),
                                viridis_lut(0.0, 0.1), 144, 96)
    # opaque transfer function: every ray stops at the hard-coded 0.99 (volume.frag.glsl:87)
    out["opaque"] = (vol, Camera.isometric_view(distance=3.0), Light.default(), RenderConfig.balanced(),
                     viridis_lut(0.0, 1.0), 128, 128)
    # zero gradients: plateau of constant density -> normalize(vec3(0)) = NaN -> max(NaN, 0) (SURVEY a-7)
    block = np.zeros((32, 32, 32), np.float32)
    block[8:24, 8:24, 8:24] = 0.6
    out["plateau_nan_normals"] = (Volume(data=block, normals=oracle.normals(block)), Camera.isometric_view(distance=3.0),
                                  Light.directional([1, -1, 0]), RenderConfig.balanced(), viridis_lut(0.0, 0.5), 128, 128)
    # non-cubic array: moderngl hands `shape` to GL as (width, height, depth) over the C-order bytes (manager.py:95-97),
    # so a (20, 28, 36) array is sampled as the same bytes viewed [depth=36][height=28][width=20].  The blob is
    # smooth IN THAT VIEW (a blob that is smooth in numpy order turns into one-texel stripes, where the last bit
    # of a texture coordinate decides the colour of a pixel)
    g = [np.exp(-np.linspace(-2.5, 2.5, k) ** 2) for k in (36, 28, 20)]
    gl_view = (g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :]).astype(np.float32)
    nc = gl_view.reshape(-1).reshape(20, 28, 36)
    nc_n = np.stack(np.gradient(gl_view), axis=-1).astype(np.float32).reshape(20, 28, 36, 3)
    out["non_cubic"] = (Volume(data=nc, normals=nc_n), Camera.isometric_view(distance=3.0), Light.default(),
                        RenderConfig.balanced(), viridis_lut(0.0, 0.6), 128, 96)
    # LUT sizes other than 256 (set_transfer_functions(size=...)); opacity 0 at density 0, like every BASELINE config
    for size in (2, 17, 1024):
        out[f"lut_{size}"] = (vol, Camera.isometric_view(distance=3.0), light, RenderConfig.fast(),
                              viridis_lut(0.0, 0.4, size), 96, 96)
    # Opacity > 0 at density 0: the first sample of every ray sits exactly ON the box surface, and whether it passes
    # the shader's inclusive [0,1] test is decided by the last bit of the ray/box arithmetic -- a rounding coin flip
    # that no two GL implementations resolve alike.  Kept as a documented limit of parity (looser assertion).
    out["surface_coin_flip"] = (vol, Camera.isometric_view(distance=3.0), light, RenderConfig.fast(),
                                viridis_lut(0.05, 0.4), 96, 96)
    if small:
        keep = ("c1_iso", "c1_rolled", "no_normals", "pm1_fast", "pm1_high_quality", "opaque", "plateau_nan_normals",
                "non_cubic", "lut_17")
        out = {k: out[k] for k in keep}
    return out
