"""Pins the CPU oracle: golden normals from the reference (bit-exact), an independent numpy
restatement of the shader, closed-form known answers, and the survey's probe counts for C1.

Pixels remain "parity unpinned" against a real OpenGL run (none is possible here); see
oracle/pyvr_oracle.c.
"""

import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                       Volume, build_rgba_lut, create_sample_volume)

from scenes import c1_scene, image_metrics, viridis_lut
import shader_numpy


def test_normals_bit_exact_vs_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "normals.npz"))
    cases = [k[:-4] for k in z.files if k.endswith("__in")]
    assert len(cases) >= 7
    for name in cases:
        got = oracle.normals(z[name + "__in"])
        want = z[name + "__out"]
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    big = oracle.normals(create_sample_volume(128, "double_sphere"))
    assert hashlib.sha256(big.tobytes()).hexdigest() == meta["normal_volume_sha256"]["double_sphere_128"]


def test_shader_source_is_the_one_restated(golden_dir):
    """If /root/reference is mounted (build container), the shader must still hash to what the oracle restates."""
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    path = "/root/reference/pyvr/shaders/volume.frag.glsl"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted on this box")
    text = open(path, "rb").read()
    assert hashlib.sha256(text).hexdigest() == meta["shader_sha256"]["pyvr/shaders/volume.frag.glsl"]
    src = text.decode()
    for needle in ("accumulated_alpha < 0.99", "exp(-alpha_tf * step_size / reference_step_size)",
                   "vec3(tex_coord.z, tex_coord.y, tex_coord.x)", "ambient_light + diffuse_light * diffuse_intensity"):
        assert needle in src, needle


@pytest.mark.parametrize("with_normals", [True, False])
def test_oracle_matches_independent_numpy_restatement(with_normals):
    data = create_sample_volume(24, "double_sphere")
    vol = Volume(data=data, normals=oracle.normals(data) if with_normals else None)
    cam = Camera(azimuth=0.7, elevation=0.35, roll=0.2, distance=2.5)
    light = Light.directional([1, -1, 0.5])
    cfg = RenderConfig(step_size=0.02, max_steps=120)
    lut = viridis_lut(0.0, 0.6)
    w, h = 56, 40
    img, acc, st = oracle.render(vol, cam, light, cfg, lut, w, h, want_accum=True)
    img2, acc2, samples2 = shader_numpy.render(vol, cam, light, cfg, lut, w, h)
    assert abs(st["samples"] - samples2) <= 2e-3 * samples2
    m = image_metrics(img, img2)
    assert m["max_abs"] <= 1 and m["frac_identical"] > 0.99, m
    # silhouette pixels may differ by one sample; everything else agrees to float32 round-off
    close = np.isclose(acc, acc2, rtol=1e-4, atol=2e-5).all(axis=-1)
    assert close.mean() > 0.995


def test_c1_work_counts_match_survey_probe():
    """SURVEY.md section 8 d / appendix A: 73 930 rays hit, ~4.34 M samples, 1 667 early stops."""
    data = create_sample_volume(128, "double_sphere")
    vol, light, lut = c1_scene(128, normals=oracle.normals(data))
    img, _, st = oracle.render(vol, Camera.isometric_view(distance=3.0), light, RenderConfig.balanced(), lut, 512, 512)
    assert st["rays_hit"] == 73930
    assert st["rays_terminated"] == 1667
    assert abs(st["samples"] - 4342927) < 1000
    assert img.shape == (512, 512, 4) and img[..., 3].max() >= 250


def _homogeneous(density, alpha_tf, cfg, light=None, size=8, width=24, height=24):
    """Constant-density cube, constant LUT: every in-box sample adds the same alpha."""
    vol = Volume(data=np.full((size, size, size), density, np.float32))
    lut = np.zeros((16, 4), np.float32)
    lut[:, :3] = (0.5, 0.25, 1.0)
    lut[:, 3] = alpha_tf
    light = light or Light.ambient_only(1.0)
    cam = Camera.front_view(distance=3.0)
    return oracle.render(vol, cam, light, cfg, lut, width, height, want_accum=True), lut


def test_homogeneous_medium_known_answer():
    """acc_a after n samples of alpha a: 1-(1-a)^n with a = 1-exp(-alpha_tf*step/ref) (volume.frag.glsl:101,115)."""
    cfg = RenderConfig(step_size=0.05, max_steps=50, reference_step_size=0.01)
    (img, acc, st), _ = _homogeneous(0.5, 0.04, cfg)
    centre = acc[12, 12]
    a = 1.0 - np.exp(-0.04 * 0.05 / 0.01)
    # the central ray crosses the unit cube front to back: thickness 1.0 -> 20 or 21 samples
    n = np.log1p(-centre[3]) / np.log1p(-a)
    assert abs(n - round(n)) < 1e-3 and round(n) in (20, 21)
    # colour = rgb * light(=1) * accumulated alpha; the normal is the zero vector -> NaN -> diffuse 0
    np.testing.assert_allclose(centre[:3], np.array([0.5, 0.25, 1.0]) * centre[3], rtol=1e-5)
    # blended bytes: (C*A, A*A), round to nearest
    want = np.rint(np.array([*(centre[:3] * centre[3]), centre[3] ** 2]) * 255)
    np.testing.assert_array_equal(img[12, 12], want.astype(np.uint8))
    # corner pixels miss the box: cleared frame
    assert not img[0, 0].any() and not acc[0, 0].any()


def test_opacity_correction_is_step_size_invariant():
    """Beer-Lambert consistency (reference tests/test_config_opacity_correction.py:179-206): the same TF
    accumulates (almost) the same alpha at any step size because alpha' = 1-exp(-alpha*dt/ref)."""
    alphas = []
    for step, steps in ((0.02, 200), (0.01, 400), (0.005, 800), (0.0025, 1600)):
        (img, acc, st), _ = _homogeneous(0.5, 0.05, RenderConfig(step_size=step, max_steps=steps))
        alphas.append(float(acc[12, 12, 3]))
    want = 1.0 - np.exp(-0.05 * 1.0 / 0.01)
    for a in alphas:  # thickness 1.0 +- one step
        assert abs(a - want) < 0.02, alphas
    assert max(alphas) - min(alphas) < 0.02


def test_zero_alpha_stays_zero_and_termination_constant():
    cfg = RenderConfig(step_size=0.01, max_steps=500)
    (img, acc, st), _ = _homogeneous(0.5, 0.0, cfg)
    assert not img.any() and st["rays_terminated"] == 0 and st["samples"] > 0
    # opaque medium: rays stop at the first sample count reaching >= 0.99, regardless of RenderConfig.opacity_threshold
    (img, acc, st), _ = _homogeneous(0.5, 1.0, RenderConfig(step_size=0.01, max_steps=500, opacity_threshold=0.5))
    a = 1.0 - np.exp(-1.0)
    n_stop = int(np.ceil(np.log(0.01) / np.log(1 - a)))
    assert abs(float(acc[12, 12, 3]) - (1 - (1 - a) ** n_stop)) < 1e-5
    assert st["rays_terminated"] == st["rays_hit"]


def test_loop_bound_is_max_steps_not_t_far():
    """'fast' preset reach is 0.02*100 = 2.0 < chord through a +-1 cube seen from distance 3 along a diagonal."""
    vol = Volume(data=np.full((8, 8, 8), 0.5, np.float32),
                 min_bounds=np.array([-1, -1, -1], np.float32), max_bounds=np.array([1, 1, 1], np.float32))
    lut = np.zeros((4, 4), np.float32)
    lut[:, 3] = 0.001
    cam = Camera.front_view(distance=3.0)
    _, acc, st = oracle.render(vol, cam, Light.ambient_only(1.0), RenderConfig(step_size=0.02, max_steps=50),
                               lut, 9, 9, want_accum=True)
    a = 1 - np.exp(-0.001 * 0.02 / 0.01)
    # the central ray enters at x=1 and would need 100 steps to cross; only 50 are taken
    assert abs(acc[4, 4, 3] - (1 - (1 - a) ** 50)) < 1e-6


def test_rows_are_bottom_up_and_axes_follow_the_swizzle():
    """Bright voxel block at high numpy-axis-2 (world z): seen from the front (+x, up = +z) it must be in
    the TOP of the picture, i.e. in the LAST rows of the returned buffer (row 0 = bottom)."""
    data = np.zeros((16, 16, 16), np.float32)
    data[6:10, 6:10, 12:16] = 1.0
    vol = Volume(data=data)
    lut = np.zeros((8, 4), np.float32)
    lut[:, :3] = 1.0
    lut[:, 3] = np.linspace(0, 1, 8)
    img, _, _ = oracle.render(vol, Camera.front_view(distance=3.0), Light.ambient_only(1.0),
                              RenderConfig.balanced(), lut, 32, 32)
    rows = img[..., 3].sum(axis=1).astype(float)
    assert rows[16:].sum() > 10 * max(rows[:16].sum(), 1)
    # same block seen from the side (+y): world x (numpy axis 0) is horizontal; block is centred -> symmetric
    data2 = np.zeros((16, 16, 16), np.float32)
    data2[12:16, 6:10, 6:10] = 1.0   # high world x
    img2, _, _ = oracle.render(Volume(data=data2), Camera.side_view(distance=3.0), Light.ambient_only(1.0),
                               RenderConfig.balanced(), lut, 32, 32)
    cols = img2[..., 3].sum(axis=0).astype(float)
    assert abs(cols[:16].sum() - cols[16:].sum()) > 0.5 * cols.sum()  # off-centre horizontally


def test_non_cubic_volume_uses_gl_width_height_depth_addressing():
    """texture3d(shape) means (w,h,d) = shape over C-order bytes (manager.py:95-97): for a (4,6,8) array
    the oracle must equal rendering the re-viewed (8,6,4) cubic-convention array."""
    # What GL sees for a (4, 6, 8) array: width 4 (fastest), height 6, depth 8, i.e. the same bytes viewed
    # as [depth=8, height=6, width=4].  Build a blob that is smooth and ~0 at the faces IN THAT VIEW
    # (whether the entry sample exactly on the box surface counts as inside is a rounding coin-flip that
    # only matters where the data is non-zero on the faces), then hand the oracle the (4, 6, 8) array.
    g = [np.exp(-np.linspace(-2.5, 2.5, n) ** 2) for n in (8, 6, 4)]
    gl_view = (g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :]
               * np.linspace(0.6, 1.0, 4)).astype(np.float32)
    data = gl_view.reshape(-1).reshape(4, 6, 8)
    lut = viridis_lut(0.0, 0.6)
    args = (Camera.isometric_view(distance=3.0), Light.default(), RenderConfig.fast(), lut, 40, 40)
    img_a, _, st = oracle.render(Volume(data=data), *args)
    img_b, _, _ = shader_numpy.render(Volume(data=data), *args)
    assert image_metrics(img_a, img_b)["max_abs"] <= 1
    # world x spans the depth axis (8 texels), world z the width axis (4 texels): the cubic-convention
    # volume gl_view[ix, iy, iz] rendered directly must give the same picture
    img_c, _, st_c = oracle.render(Volume(data=np.ascontiguousarray(gl_view)), *args)
    assert st_c["rays_hit"] == st["rays_hit"]
    assert img_a[..., 3].max() > 100


def test_row_subset_matches_full_render():
    data = create_sample_volume(32, "torus")
    vol = Volume(data=data, normals=oracle.normals(data))
    args = (vol, Camera.isometric_view(distance=3.0), Light.default(), RenderConfig.balanced(), viridis_lut(), 64, 48)
    full, _, st_full = oracle.render(*args)
    part, _, st_part = oracle.render(*args, rows=(3, 48, 8))
    np.testing.assert_array_equal(part[3::8], full[3::8])
    assert not part[4].any() and 0 < st_part["samples"] < st_full["samples"]
