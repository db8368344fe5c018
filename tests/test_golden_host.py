"""Host mirror classes vs fixtures generated from the reference's own Python (tests/golden/make_golden.py)."""

import hashlib
import json
import os

import numpy as np
import pytest

from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                       build_rgba_lut, create_sample_volume)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_camera_vectors_and_matrices(golden_dir):
    records = json.load(open(os.path.join(golden_dir, "camera.json")))
    assert len(records) >= 18
    for rec in records:
        cam = Camera.from_dict(rec["params"])
        pos, up = cam.get_camera_vectors()
        np.testing.assert_allclose(pos, rec["position"], rtol=0, atol=1e-12, err_msg=rec["name"])
        np.testing.assert_allclose(up, rec["up"], rtol=0, atol=1e-12, err_msg=rec["name"])
        view = cam.get_view_matrix()
        assert view.dtype == np.float32 and view.shape == (4, 4)
        np.testing.assert_allclose(view, np.array(rec["view"]), rtol=0, atol=1e-7, err_msg=rec["name"])
        for aspect, want in rec["proj"].items():
            proj = cam.get_projection_matrix(float(aspect))
            assert proj.dtype == np.float32
            np.testing.assert_array_equal(proj, np.array(want, dtype=np.float32), err_msg=rec["name"])


def test_camera_known_answers():
    # reference tests/test_camera/test_control.py:32-33 and SURVEY.md appendix A
    from pyvr_b200 import get_camera_pos

    pos, up = get_camera_pos(np.array([0.0, 0.0, 0.0]), 0.0, 0.0, 0.0, 3.0)
    np.testing.assert_allclose(pos, [0, 0, 3], atol=1e-6)
    np.testing.assert_allclose(up, [0, 1, 0], atol=1e-6)
    pos, up = Camera.front_view(distance=3.0).get_camera_vectors()
    np.testing.assert_allclose(pos, [3, 0, 0], atol=1e-6)
    np.testing.assert_allclose(up, [0, 0, 1], atol=1e-6)
    proj = Camera.isometric_view().get_projection_matrix(1.0)
    assert proj[3, 3] == 0 and abs(proj[0, 0] - 2.4142137) < 1e-6 and proj[3, 2] == -1


def test_camera_validation_errors():
    with pytest.raises(ValueError, match="distance must be positive"):
        Camera(distance=0)
    with pytest.raises(ValueError, match="target must be a 3D numpy array"):
        Camera(target=[0, 0, 0])
    with pytest.raises(ValueError, match="fov must be between"):
        Camera(fov=4.0)


def test_lights(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "lights.json")))
    linked = Light.camera_linked()
    linked.update_from_camera(Camera.isometric_view(distance=3.0))
    got = {"default": Light.default(), "directional_1m10": Light.directional([1, -1, 0]),
           "directional_custom": Light.directional(np.array([0.2, 0.5, -1.0]), ambient=0.3, diffuse=0.6, distance=4.0),
           "point": Light.point_light([5, 5, 5], ambient=0.1, diffuse=0.7),
           "ambient_only": Light.ambient_only(0.3), "camera_linked_iso": linked}
    assert set(got) == set(want)
    for name, light in got.items():
        w = want[name]
        np.testing.assert_allclose(light.position, w["position"], atol=1e-7, err_msg=name)
        np.testing.assert_allclose(light.target, w["target"], atol=1e-7, err_msg=name)
        assert light.ambient_intensity == w["ambient"] and light.diffuse_intensity == w["diffuse"]
        np.testing.assert_allclose(light.get_direction(), w["direction"], atol=1e-7, err_msg=name)
    with pytest.raises(ValueError, match="not linked"):
        Light.default().update_from_camera(Camera())
    with pytest.raises(ValueError, match="ambient_intensity"):
        Light(ambient_intensity=1.5)


def test_presets(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "presets.json")))
    got = {"preview": RenderConfig.preview(), "fast": RenderConfig.fast(), "balanced": RenderConfig.balanced(),
           "high_quality": RenderConfig.high_quality(), "ultra_quality": RenderConfig.ultra_quality(),
           "default": RenderConfig()}
    for name, cfg in got.items():
        w = want[name]
        assert (cfg.step_size, cfg.max_steps, cfg.early_ray_termination, cfg.opacity_threshold,
                cfg.reference_step_size) == (w["step_size"], w["max_steps"], w["early_ray_termination"],
                                             w["opacity_threshold"], w["reference_step_size"]), name
        assert cfg.estimate_samples_per_ray() == w["samples_per_ray"]
        assert cfg.estimate_render_time_relative() == pytest.approx(w["relative_time"])
        assert repr(cfg) == w["repr"]
    with pytest.raises(ValueError, match="step_size must be positive"):
        RenderConfig(step_size=0)
    with pytest.raises(ValueError, match="max_steps must be at least 1"):
        RenderConfig(max_steps=0)
    with pytest.raises(ValueError, match="opacity_threshold"):
        RenderConfig(opacity_threshold=1.5)
    assert RenderConfig.balanced().with_step_size(0.005).reference_step_size == 0.01
    assert RenderConfig.balanced().with_max_steps(600).max_steps == 600


def test_luts(golden_dir):
    z = np.load(os.path.join(golden_dir, "luts.npz"))
    pts = [(0.0, (0.0, 0.0, 0.2)), (0.25, (0.1, 0.9, 0.3)), (0.6, (1.0, 0.5, 0.0)), (1.0, (1.0, 1.0, 1.0))]
    got = {
        "otf_linear_0_1": OpacityTransferFunction.linear(0.0, 1.0).to_lut(),
        "otf_linear_0_0p3": OpacityTransferFunction.linear(0.0, 0.3).to_lut(),
        "otf_linear_0_0p1_64": OpacityTransferFunction.linear(0.0, 0.1).to_lut(64),
        "otf_one_step": OpacityTransferFunction.one_step(0.5, 0.0, 0.8).to_lut(),
        "otf_peaks": OpacityTransferFunction.peaks([0.3, 0.7], opacity=0.9, eps=0.05, base=0.1).to_lut(),
        "otf_custom": OpacityTransferFunction([(0.0, 0.0), (0.3, 0.1), (0.8, 0.9), (1.0, 0.5)]).to_lut(100),
        "ctf_gray": ColorTransferFunction.grayscale().to_lut(),
        "ctf_custom": ColorTransferFunction(pts).to_lut(),
        "ctf_custom_17": ColorTransferFunction(pts).to_lut(17),
        "ctf_single": ColorTransferFunction.single_color((0.2, 0.4, 0.6)).to_lut(32),
    }
    assert set(got) == set(z.files)
    for name, lut in got.items():
        assert lut.dtype == np.float32
        np.testing.assert_array_equal(lut, z[name], err_msg=name)


def test_rgba_lut_packing():
    ctf = ColorTransferFunction.grayscale(lut_size=64)
    otf = OpacityTransferFunction.linear(0.0, 0.5, lut_size=128)
    lut = build_rgba_lut(ctf, otf)
    assert lut.shape == (128, 4) and lut.dtype == np.float32       # size = max of the two (manager.py:163-166)
    np.testing.assert_array_equal(lut[:, :3], ctf.to_lut(128))
    np.testing.assert_array_equal(lut[:, 3], otf.to_lut(128))
    assert build_rgba_lut(ctf, otf, 16).shape == (16, 4)


def test_colormap_fallback_tables():
    ctf = ColorTransferFunction.from_colormap("viridis")
    lut = ctf.to_lut()
    np.testing.assert_allclose(lut[0] * 255, [68, 1, 84], atol=0.51)
    np.testing.assert_allclose(lut[-1] * 255, [253, 231, 37], atol=0.51)
    assert len(ctf.control_points) == 256


def test_sample_volumes(golden_dir):
    z = np.load(os.path.join(golden_dir, "volumes.npz"))
    for shape in z.files:
        got = create_sample_volume(16, shape)
        assert got.dtype == np.float32 and got.shape == (16, 16, 16)
        np.testing.assert_array_equal(got, z[shape], err_msg=shape)
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    for key, want in meta["sample_volume_sha256"].items():
        shape, n = key.rsplit("_", 1)
        assert _sha(create_sample_volume(int(n), shape)) == want, key
    with pytest.raises(ValueError, match="Unknown shape"):
        create_sample_volume(8, "teapot")
    # trap 8 of SURVEY.md: 'xy' meshgrid puts the analytic x offset along numpy axis 1
    v = create_sample_volume(64, "double_sphere")
    assert np.unravel_index(np.argmax(v), v.shape) == (31, 22, 31)
