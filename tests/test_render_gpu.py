"""Parity of the CUDA march (through VolumeRenderer / the C ABI) against the CPU oracle.

Tolerance (BASELINE.json north_star): per-channel |delta| <= 2/255 on >= 99.9 % of pixels and
PSNR >= 45 dB.  The STRICT mode is held to a much tighter bar: it follows the oracle statement by
statement, so only expf round-off separates them.
"""

import os

import numpy as np
import pytest

import oracle
from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                       Volume, build_rgba_lut, compute_normal_volume, create_sample_volume)
from pyvr_b200.cuda_renderer import VolumeRenderer

from scenes import assert_parity, c1_scene, image_metrics, turntable_camera, viridis_lut

pytestmark = pytest.mark.gpu


def _frame(r):
    return np.frombuffer(r.render(), dtype=np.uint8).reshape(r.height, r.width, 4)


def _render_gpu(vol, cam, light, cfg, lut, w, h, **kw):
    with VolumeRenderer(w, h, config=cfg, light=light, **kw) as r:
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_lut(lut)
        img = _frame(r).copy()
        return img, r.stats


@pytest.fixture(scope="module")
def c1():
    data = create_sample_volume(128, "double_sphere")
    vol, light, lut = c1_scene(128, normals=oracle.normals(data))
    return vol, light, lut


CAMERAS = {
    "iso": lambda: Camera.isometric_view(distance=3.0),
    "front": lambda: Camera.front_view(distance=3.0),
    "side": lambda: Camera.side_view(distance=3.0),
    "top": lambda: Camera.top_view(distance=3.0),
    "rolled": lambda: Camera(azimuth=1.1, elevation=-0.4, roll=0.7, distance=2.2),
    "inside": lambda: Camera(azimuth=0.3, elevation=0.2, distance=0.3),   # camera inside the box: t_near clamps to 0
}


@pytest.mark.parametrize("cam_name", list(CAMERAS))
def test_c1_strict_matches_oracle_almost_bit_for_bit(c1, cam_name):
    vol, light, lut = c1
    cam, cfg = CAMERAS[cam_name](), RenderConfig.balanced()
    want, _, counters = oracle.render(vol, cam, light, cfg, lut, 512, 512)
    got, stats = _render_gpu(vol, cam, light, cfg, lut, 512, 512, strict=True)
    m = image_metrics(got, want)
    assert m["max_abs"] <= 1 and m["frac_identical"] >= 0.9999, m
    # a ray whose accumulated alpha sits within an ulp of 0.99 may stop one sample apart (expf round-off)
    assert abs(stats["samples"] - counters["samples"]) <= 16
    assert stats["rays_hit"] == counters["rays_hit"]
    assert abs(stats["rays_terminated"] - counters["rays_terminated"]) <= 2


@pytest.mark.parametrize("cam_name", list(CAMERAS))
@pytest.mark.parametrize("ess", [False, True])
def test_c1_fast_within_tolerance(c1, cam_name, ess):
    vol, light, lut = c1
    cam, cfg = CAMERAS[cam_name](), RenderConfig.balanced()
    want, _, counters = oracle.render(vol, cam, light, cfg, lut, 512, 512)
    got, stats = _render_gpu(vol, cam, light, cfg, lut, 512, 512, empty_space_skipping=ess)
    m = assert_parity(got, want)
    assert m["max_abs"] <= 2, m
    assert stats["rays_hit"] == counters["rays_hit"]
    assert abs(stats["samples"] - counters["samples"]) <= 2e-4 * counters["samples"] + 64
    if ess:
        assert stats["samples_fetched"] < 0.9 * stats["samples"]     # 79 % of this volume maps to alpha 0
    else:
        assert stats["samples_fetched"] == stats["samples"]


def test_c1_counts_match_survey_probe(c1):
    vol, light, lut = c1
    _, stats = _render_gpu(vol, Camera.isometric_view(distance=3.0), light, RenderConfig.balanced(), lut,
                           512, 512, strict=True)
    assert stats["rays_hit"] == 73930 and stats["rays_terminated"] == 1667
    assert abs(stats["samples"] - 4342927) < 1000


def test_ess_is_exact(c1):
    """Skipping only drops samples whose contribution is exactly +0: pre-blend floats identical."""
    vol, light, lut = c1
    out = {}
    for ess in (False, True):
        with VolumeRenderer(384, 256, config=RenderConfig.high_quality(), light=light, empty_space_skipping=ess) as r:
            r.load_volume(vol)
            r.set_camera(turntable_camera(37))
            r.set_lut(lut)
            out[ess] = r.render_accum()
    assert np.array_equal(out[False].view(np.uint32), out[True].view(np.uint32))


@pytest.mark.parametrize("preset", ["preview", "fast", "balanced", "high_quality", "ultra_quality"])
def test_presets_bounds_pm1(preset):
    """examples/benchmark.py scene (bounds +-1, default light, linear(0,0.1)): fast/ultra truncate long chords."""
    data = create_sample_volume(64, "double_sphere")
    vol = Volume(data=data, normals=oracle.normals(data),
                 min_bounds=np.array([-1, -1, -1], np.float32), max_bounds=np.array([1, 1, 1], np.float32))
    cfg = getattr(RenderConfig, preset)()
    lut = viridis_lut(0.0, 0.1)
    cam = Camera.from_spherical(target=np.array([0.0, 0.0, 0.0]), azimuth=np.pi / 4, elevation=np.pi / 6,
                                roll=0.0, distance=3.0)
    want, _, counters = oracle.render(vol, cam, Light.default(), cfg, lut, 200, 160)
    for kw in ({"strict": True}, {}):
        got, stats = _render_gpu(vol, cam, Light.default(), cfg, lut, 200, 160, **kw)
        m = assert_parity(got, want)
        assert stats["rays_hit"] == counters["rays_hit"]
        if kw:
            assert m["max_abs"] <= 1 and abs(stats["samples"] - counters["samples"]) <= 16


def test_volume_without_normals_uses_density_as_normal():
    data = create_sample_volume(48, "torus")
    vol = Volume(data=data)                      # normals None: shader samples the scalar texture as the normal
    cam, light, cfg, lut = Camera.isometric_view(distance=2.5), Light.directional([1, 0, 0]), RenderConfig.balanced(), viridis_lut(0, 0.5)
    want, _, _ = oracle.render(vol, cam, light, cfg, lut, 160, 160)
    got, _ = _render_gpu(vol, cam, light, cfg, lut, 160, 160, strict=True)
    assert image_metrics(got, want)["max_abs"] <= 1
    got, _ = _render_gpu(vol, cam, light, cfg, lut, 160, 160)
    assert_parity(got, want)
    # and it differs from ambient-only shading, i.e. the (1,0,0) normal really is lit
    dark, _ = _render_gpu(vol, cam, Light.directional([-1, 0, 0]), cfg, lut, 160, 160)
    assert np.abs(dark.astype(int) - got.astype(int)).max() > 5


def test_zero_gradient_voxels_take_the_nan_path():
    """Exactly-zero normals (cube shape, VTK backgrounds): normalize(0) = NaN -> diffuse 0, ambient only."""
    data = create_sample_volume(40, "cube")
    vol = Volume(data=data, normals=oracle.normals(data))
    lut = np.zeros((64, 4), np.float32)
    lut[:, :3] = 0.8
    lut[:, 3] = 0.2                              # opaque even where density == 0 and the normal is the zero vector
    args = (Camera.isometric_view(distance=3.0), Light.default(), RenderConfig.balanced(), lut, 128, 128)
    want, _, _ = oracle.render(vol, *args)
    got, _ = _render_gpu(vol, *args)
    assert_parity(got, want)
    assert not np.isnan(want.astype(float)).any() and got[..., 3].max() > 200


def test_non_cubic_volume():
    g = [np.exp(-np.linspace(-2.5, 2.5, n) ** 2) for n in (40, 24, 12)]
    gl_view = (g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :]).astype(np.float32)
    data = np.ascontiguousarray(gl_view.reshape(-1).reshape(12, 24, 40))
    vol = Volume(data=data, normals=oracle.normals(data),
                 min_bounds=np.array([-1, -0.6, -0.3], np.float32), max_bounds=np.array([1, 0.6, 0.3], np.float32))
    args = (Camera.isometric_view(distance=3.0), Light.default(), RenderConfig.balanced(), viridis_lut(0, 0.8), 192, 128)
    want, _, counters = oracle.render(vol, *args)
    got, stats = _render_gpu(vol, *args, strict=True)
    assert image_metrics(got, want)["max_abs"] <= 1 and abs(stats["samples"] - counters["samples"]) <= 16
    for layout in ("linear", "swizzle"):
        os.environ["PYVR_CUDA_LAYOUT"] = layout
        try:
            got, _ = _render_gpu(vol, *args)
        finally:
            del os.environ["PYVR_CUDA_LAYOUT"]
        assert_parity(got, want)


@pytest.mark.parametrize("size", [1, 2, 17, 256, 1024, 3000, 4096])   # 3000: static + dynamic shared memory just above 48 KiB
def test_lut_sizes(size):
    data = create_sample_volume(32, "sphere")
    vol = Volume(data=data, normals=oracle.normals(data))
    lut = build_rgba_lut(ColorTransferFunction.from_colormap("plasma"), OpacityTransferFunction.linear(0.05, 0.4), size)
    args = (Camera.isometric_view(distance=3.0), Light.default(), RenderConfig.balanced(), lut, 96, 96)
    want, _, _ = oracle.render(vol, *args)
    got, _ = _render_gpu(vol, *args)
    assert_parity(got, want)


@pytest.mark.parametrize("layout", ["linear", "swizzle"])
def test_half_texels_within_tolerance(c1, layout):
    vol, light, lut = c1
    cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
    want, _, _ = oracle.render(vol, cam, light, cfg, lut, 512, 512)
    os.environ["PYVR_CUDA_LAYOUT"] = layout
    try:
        got, _ = _render_gpu(vol, cam, light, cfg, lut, 512, 512, texel_format="f16")
    finally:
        del os.environ["PYVR_CUDA_LAYOUT"]
    assert_parity(got, want)


@pytest.mark.parametrize("layout", ["swizzle"])
def test_layouts_are_bit_identical(c1, layout):
    """The texel layout only moves bytes around: every layout must give the same float image."""
    vol, light, lut = c1
    out = {}
    for name in ("linear", layout):
        os.environ["PYVR_CUDA_LAYOUT"] = name
        try:
            with VolumeRenderer(200, 120, config=RenderConfig.balanced(), light=light) as r:
                r.load_volume(vol)
                r.set_camera(turntable_camera(123))
                r.set_lut(lut)
                out[name] = r.render_accum()
        finally:
            del os.environ["PYVR_CUDA_LAYOUT"]
    assert np.array_equal(out["linear"].view(np.uint32), out[layout].view(np.uint32))


@pytest.mark.parametrize("texel_format", ["f32", "f16"])
@pytest.mark.parametrize("shape", [(64, 64, 64), (33, 47, 58)])
def test_brick8_layout_is_bit_identical(texel_format, shape):
    """2x2x2-texel bricks (the layout for sparse rays, option "brick8") hold the same texels as the row layouts:
    fast and STRICT marches give the same float image, odd sizes and the apron included."""
    rng = np.random.default_rng(5)
    data = rng.random(shape, dtype=np.float32)
    normals = rng.standard_normal(shape + (3,)).astype(np.float32)
    vol = Volume(data=data, normals=normals)
    lut = build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.2))
    for strict in (False, True):
        out = {}
        for b8 in ("0", "1", "1+multi"):        # "1+multi": bricks marched with several samples in flight (f16 only)
            os.environ["PYVR_CUDA_BRICK8"] = b8[0]
            os.environ["PYVR_CUDA_TWO_SAMPLES"] = "1" if b8.endswith("multi") else "0"
            try:
                with VolumeRenderer(160, 96, config=RenderConfig.balanced(), light=Light.directional([1, -1, 0]),
                                    texel_format=texel_format, strict=strict) as r:
                    r.load_volume(vol)
                    r.set_camera(turntable_camera(77))
                    r.set_lut(lut)
                    out[b8] = r.render_accum()
            finally:
                del os.environ["PYVR_CUDA_BRICK8"], os.environ["PYVR_CUDA_TWO_SAMPLES"]
        assert np.array_equal(out["0"].view(np.uint32), out["1"].view(np.uint32)), (texel_format, shape, strict)
        assert np.array_equal(out["0"].view(np.uint32), out["1+multi"].view(np.uint32)), (texel_format, shape, strict)


@pytest.mark.parametrize("ess", [False, True])
def test_two_samples_per_iteration_is_bit_identical(c1, ess):
    """f16x4 z-pair march with two samples of a run in flight per ray (option "two_samples", automatic for large
    volumes): same samples, same order, the second one dropped when the first saturates the ray -- same floats.
    The opaque step transfer function makes many rays stop early, on odd and even sample indices."""
    vol, light, _ = c1
    luts = [viridis_lut(0.0, 0.3),
            build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.one_step(0.3, 0.0, 0.9))]
    for lut in luts:
        out = {}
        for two in ("0", "1"):
            os.environ["PYVR_CUDA_TWO_SAMPLES"] = two
            try:
                with VolumeRenderer(320, 200, config=RenderConfig.high_quality(), light=light, texel_format="f16",
                                    empty_space_skipping=ess) as r:
                    r.load_volume(vol)
                    r.set_camera(turntable_camera(211))
                    r.set_lut(lut)
                    out[two] = (r.render_accum(), r.stats)
            finally:
                del os.environ["PYVR_CUDA_TWO_SAMPLES"]
        assert np.array_equal(out["0"][0].view(np.uint32), out["1"][0].view(np.uint32))
        for key in ("samples", "rays_hit", "rays_terminated"):
            assert out["0"][1][key] == out["1"][1][key], key


def test_render_without_volume_or_camera_returns_cleared_frame():
    with VolumeRenderer(64, 48) as r:
        assert r.render() == bytes(64 * 48 * 4)          # nothing loaded (reference tolerates this)
        r.set_camera(Camera.front_view())
        assert r.render() == bytes(64 * 48 * 4)
        data = create_sample_volume(16, "sphere")
        r.load_volume(Volume(data=data))
        assert r.render() == bytes(64 * 48 * 4)          # still no LUT
        r.set_transfer_functions(ColorTransferFunction.grayscale(), OpacityTransferFunction.linear(0, 1))
        frame = r.render()
        assert isinstance(frame, bytes) and len(frame) == 64 * 48 * 4 and any(frame)
        assert r.render_to_pil().size == (64, 48)


def test_api_errors_and_accessors():
    with VolumeRenderer(32, 32) as r:
        with pytest.raises(TypeError, match="Expected Volume instance"):
            r.load_volume(np.zeros((4, 4, 4), np.float32))
        with pytest.raises(TypeError, match="Expected Camera instance"):
            r.set_camera("front")
        with pytest.raises(TypeError, match="Expected Light instance"):
            r.set_light(None)
        with pytest.raises(TypeError, match="Expected RenderConfig instance"):
            r.set_config({"step_size": 0.1})
        assert r.get_volume() is None and r.get_camera() is None
        cfg = RenderConfig.fast()
        r.set_config(cfg)
        assert r.get_config() is cfg and isinstance(r.get_light(), Light)
        vol = Volume(data=np.zeros((4, 4, 4), np.float64))          # non-f32 data is cast (manager.py:91-92)
        r.load_volume(vol)
        assert r.get_volume() is vol


def test_batch_equals_single_and_rows_are_bottom_up(c1):
    vol, light, lut = c1
    cams = [turntable_camera(k) for k in (0, 90, 200)]
    with VolumeRenderer(256, 128, config=RenderConfig.balanced(), light=light) as r:
        r.load_volume(vol)
        r.set_lut(lut)
        batch = r.render_batch(cams)
        assert batch.shape == (3, 128, 256, 4)
        for i, cam in enumerate(cams):
            r.set_camera(cam)
            assert np.array_equal(batch[i], _frame(r))
        assert r.stats["views"] == 1
    # bottom-up: a blob high in world z (camera up) must land in the last rows
    data = np.zeros((16, 16, 16), np.float32)
    data[6:10, 6:10, 12:16] = 1.0
    lut2 = np.zeros((8, 4), np.float32)
    lut2[:, :3], lut2[:, 3] = 1.0, np.linspace(0, 1, 8)
    img, _ = _render_gpu(Volume(data=data), Camera.front_view(distance=3.0), Light.ambient_only(1.0),
                         RenderConfig.balanced(), lut2, 32, 32)
    rows = img[..., 3].sum(axis=1).astype(float)
    assert rows[16:].sum() > 10 * max(rows[:16].sum(), 1)


def test_homogeneous_medium_known_answer_on_gpu():
    vol = Volume(data=np.full((8, 8, 8), 0.5, np.float32))
    lut = np.zeros((16, 4), np.float32)
    lut[:, :3], lut[:, 3] = (0.5, 0.25, 1.0), 0.04
    cfg = RenderConfig(step_size=0.05, max_steps=50, reference_step_size=0.01)
    with VolumeRenderer(24, 24, config=cfg, light=Light.ambient_only(1.0)) as r:
        r.load_volume(vol)
        r.set_camera(Camera.front_view(distance=3.0))
        r.set_lut(lut)
        acc = r.render_accum()
    a = 1.0 - np.exp(-0.04 * 0.05 / 0.01)
    n = np.log1p(-acc[12, 12, 3]) / np.log1p(-a)
    assert abs(n - round(n)) < 2e-3 and round(n) in (20, 21)
    np.testing.assert_allclose(acc[12, 12, :3], np.array([0.5, 0.25, 1.0]) * acc[12, 12, 3], rtol=1e-4)
    assert not acc[0, 0].any()


def test_honor_config_termination_is_opt_in(c1):
    vol, light, lut = c1
    cfg = RenderConfig(step_size=0.01, max_steps=500, opacity_threshold=0.5)
    cam = Camera.isometric_view(distance=3.0)
    ref_like, _ = _render_gpu(vol, cam, light, cfg, lut, 128, 128)
    want, _, _ = oracle.render(vol, cam, light, cfg, lut, 128, 128)
    assert_parity(ref_like, want)                           # default: 0.99, like the shader
    early, stats = _render_gpu(vol, cam, light, cfg, lut, 128, 128, honor_config_termination=True)
    assert early[..., 3].max() < ref_like[..., 3].max()


@pytest.mark.parametrize("name", ["fuel", "hydrogen"])
@pytest.mark.parametrize("view", ["iso", "front"])
def test_c2_vti_volumes_high_quality_camera_linked_light(tmp_path, golden_dir, name, view):
    """BASELINE config 2: the reference's example volumes through the .vti loader (bounds +-1, normals from the
    stencil kernel, huge exactly-zero background => the normalize(0) NaN path), 1024x1024, high_quality,
    camera-linked light.  The oracle renders every 4th row to stay within seconds."""
    import json

    from pyvr_b200.dataloaders import load_vtk_volume
    from vti_writer import write_vti

    vols = np.load(os.path.join(golden_dir, "vti_volumes.npz"))
    meta = json.load(open(os.path.join(golden_dir, "vti_meta.json")))
    path = write_vti(tmp_path / f"{name}.vti", {"Scalars_": vols[name].astype(np.float32).ravel()}, meta[name]["dims_xyz"])
    vol = load_vtk_volume(path)                                  # normals on the GPU (K2)
    assert vol.has_normals and np.array_equal(vol.normals, oracle.normals(vol.data))
    assert not np.isnan(vol.normals).any() and (np.abs(vol.normals).sum(axis=-1) == 0).mean() > 0.5
    cam = Camera.isometric_view(distance=3.0) if view == "iso" else Camera.front_view(distance=3.0)
    light = Light.camera_linked()
    light.update_from_camera(cam)
    cfg, lut, size = RenderConfig.high_quality(), viridis_lut(0.0, 0.3), 1024
    rows = (1, size, 4)
    want, _, counters = oracle.render(vol, cam, light, cfg, lut, size, size, rows=rows)
    got, stats = _render_gpu(vol, cam, light, cfg, lut, size, size)
    assert stats["rays_hit"] > 300000 and got.any()
    assert_parity(got[rows[0]::rows[2]], want[rows[0]::rows[2]])


def test_c3_full_size_view_matches_oracle():
    """The headline configuration at BASELINE.json's full size: 512^3 double_sphere + normals, bounds +-1,
    1920x1080, high_quality, turntable view 37, viridis + linear(0, 0.1).  One whole frame against the oracle
    (about 2 s of CPU on the GPU box), plus the texture-unit alternative, plus size-independent properties:
    empty-space skipping and the z-pair layout leave the float accumulators untouched."""
    size, w, h = 512, 1920, 1080
    data = create_sample_volume(size, "double_sphere")
    normals = compute_normal_volume(data)                       # K2
    vol = Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32),
                 max_bounds=np.array([1, 1, 1], np.float32))
    light, cfg, lut = Light.directional([1, -1, 0]), RenderConfig.high_quality(), viridis_lut(0.0, 0.1)
    cam = turntable_camera(37)
    want, _, counters = oracle.render(vol, cam, light, cfg, lut, w, h)
    assert counters["samples"] > 3.0e8
    accs = {}
    for name, kw in (("default", {}), ("dense", dict(empty_space_skipping=False)), ("hwtex", dict(hardware_filtering=True)),
                     ("f16", dict(texel_format="f16"))):
        with VolumeRenderer(w, h, config=cfg, light=light, **kw) as r:
            r.load_volume(vol)
            r.set_camera(cam)
            r.set_lut(lut)
            got = np.frombuffer(r.render(), dtype=np.uint8).reshape(h, w, 4)
            m = assert_parity(got, want)
            assert abs(r.stats["samples"] - counters["samples"]) <= 1e-5 * counters["samples"], (name, r.stats, counters)
            print(name, m, r.stats["samples_fetched"] / r.stats["samples"])
            if name in ("default", "dense"):
                accs[name] = r.render_accum()
    assert np.array_equal(accs["default"], accs["dense"])
    os.environ["PYVR_CUDA_PAIR"] = "0"
    try:
        with VolumeRenderer(w, h, config=cfg, light=light) as r:
            r.load_volume(vol)
            r.set_camera(cam)
            r.set_lut(lut)
            assert np.array_equal(r.render_accum(), accs["default"])
    finally:
        del os.environ["PYVR_CUDA_PAIR"]


def test_render_tensor_is_the_same_frame_on_the_device(c1):
    import torch

    vol, light, lut = c1
    cams = [turntable_camera(k) for k in (0, 90, 200)]
    with VolumeRenderer(320, 200, config=RenderConfig.fast(), light=light) as r:
        r.load_volume(vol)
        r.set_lut(lut)
        r.set_camera(cams[1])
        host = np.frombuffer(r.render(), np.uint8).reshape(200, 320, 4)
        dev = r.render_tensor()
        assert dev.is_cuda and dev.dtype == torch.uint8 and np.array_equal(dev.cpu().numpy(), host)
        batch = r.render_tensor(cams)
        assert batch.shape == (3, 200, 320, 4) and np.array_equal(batch[1].cpu().numpy(), host)
        with pytest.raises(ValueError):
            r.render_tensor(out=torch.empty((1, 2, 3), dtype=torch.uint8, device="cuda"))


def test_many_active_intervals_refill_the_interval_table():
    """Striped volume: a ray along x crosses more active runs than the per-ray interval table holds (MAX_IV = 6),
    so the walk is resumed mid-ray.  Skipping must still be exact (bit-identical accumulators vs the dense
    march) and the frame within tolerance of the oracle."""
    n = 192
    ix = np.arange(n)
    stripes = (((ix // 12) % 2) == 0).astype(np.float32) * 0.8          # 8 active slabs of 12 voxels along x
    data = np.ascontiguousarray(np.broadcast_to(stripes[:, None, None], (n, n, n))).copy()
    data[:, : n // 4, :] = 0.0                                          # leave part of the volume empty as well
    vol = Volume(data=data, normals=oracle.normals(data))
    light, cfg = Light.directional([1, -1, 0]), RenderConfig.high_quality()
    lut = build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.one_step(0.5, 0.0, 0.05))
    w, h = 256, 192
    for cam in (Camera.front_view(distance=3.0), Camera(azimuth=0.35, elevation=0.2, distance=2.5)):
        accs, fetched = {}, {}
        for ess in (True, False):
            with VolumeRenderer(w, h, config=cfg, light=light, empty_space_skipping=ess) as r:
                r.load_volume(vol)
                r.set_camera(cam)
                r.set_lut(lut)
                accs[ess] = r.render_accum()
                fetched[ess] = r.stats["samples_fetched"]
                if ess:
                    got = np.frombuffer(r.render(), np.uint8).reshape(h, w, 4).copy()
        assert np.array_equal(accs[True], accs[False])
        assert fetched[True] < 0.8 * fetched[False]
        want, _, _ = oracle.render(vol, cam, light, cfg, lut, w, h)
        assert got.any()
        assert_parity(got, want)
