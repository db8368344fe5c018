"""Device-side synthetic volumes: pyvr_cuda_generate_volume must produce what the
host pipeline create_sample_volume -> compute_normal_volume -> load_volume uploads."""

import numpy as np
import pytest

import oracle
from pyvr_b200 import Camera, Light, RenderConfig, Volume, create_sample_volume
from pyvr_b200 import multi_gpu as mg
from pyvr_b200.cuda_renderer import VolumeRenderer

from scenes import viridis_lut

pytestmark = pytest.mark.gpu


def _ulp_diff(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


@pytest.mark.parametrize("shape", ["sphere", "torus", "double_sphere"])
@pytest.mark.parametrize("size", [33, 64, 100])
def test_generated_texels_match_the_host_pipeline(shape, size):
    want_s = create_sample_volume(size, shape)
    want_n = oracle.normals(want_s)                      # bit-exact restatement of compute_normal_volume
    with VolumeRenderer(64, 64) as r:
        ms = r.generate_volume(size, shape)
        got_s, got_n = r.read_texels((size, size, size))
    assert ms > 0
    # binary64 exp on the device and in numpy may differ in the last bit; after rounding to binary32 that
    # survives for at most a handful of voxels, and then by one ulp
    ulp = _ulp_diff(got_s, want_s)
    assert ulp.max() <= 1 and (ulp == 0).mean() >= 0.9999, (ulp.max(), (ulp == 0).mean())
    if ulp.max() == 0:
        assert np.array_equal(got_n, want_n)
    else:
        assert np.abs(got_n - want_n).max() <= 1e-5 * 20      # a 1-ulp scalar moves a tiny gradient's direction


def test_generated_bricks_are_slices_of_the_whole():
    size = 128
    with VolumeRenderer(64, 64) as r:
        r.generate_volume(size, "double_sphere")
        whole_s, whole_n = r.read_texels((size,) * 3)
        for world in (2, 8):
            for rank in range(world):
                b = mg.brick_of_rank((size,) * 3, rank, world)
                r.generate_volume(size, "double_sphere", brick=b)
                s, n = r.read_texels(b.dims)
                assert np.array_equal(s, whole_s[b.slices()]) and np.array_equal(n, whole_n[b.slices()])


@pytest.mark.parametrize("texels", ["f32", "f16"])
def test_render_from_generated_volume_equals_render_from_uploaded(texels):
    size, w, h = 128, 320, 240
    data = create_sample_volume(size, "double_sphere")
    vol = Volume(data=data, normals=oracle.normals(data))
    cam, cfg, light, lut = Camera.isometric_view(distance=3.0), RenderConfig.balanced(), Light.directional([1, -1, 0]), viridis_lut()
    frames, samples = [], []
    for generated in (False, True):
        with VolumeRenderer(w, h, config=cfg, light=light, texel_format=texels) as r:
            if generated:
                r.generate_volume(size, "double_sphere", vol.min_bounds, vol.max_bounds)
            else:
                r.load_volume(vol)
            r.set_camera(cam)
            r.set_lut(lut)
            frames.append(np.frombuffer(r.render(), np.uint8).reshape(h, w, 4).copy())
            samples.append(r.stats["samples"])
    assert samples[0] == samples[1]
    d = np.abs(frames[0].astype(int) - frames[1].astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.9999


def test_generate_volume_argument_errors():
    with VolumeRenderer(32, 32) as r:
        with pytest.raises(ValueError, match="Unknown shape"):
            r.generate_volume(32, "helix")
        with pytest.raises(RuntimeError, match="size must be at least 2"):
            r.generate_volume(1, "sphere")
