"""The CUDA march against REAL runs of the reference's shader (not only against the oracle).

Same GPU box, same process: ``oracle.gl`` executes pyvr/shaders/volume.frag.glsl verbatim on Mesa llvmpipe
(host cores) and ``VolumeRenderer`` renders the same scene through the C ABI on the B200.  Tolerance:
BASELINE.json's -- per-channel |delta| <= 2/255 on >= 99.9 % of pixels, PSNR >= 45 dB.
"""

import numpy as np
import pytest

import oracle
import oracle.gl as ogl
from pyvr_b200 import Camera, Light, RenderConfig, Volume, create_sample_volume
from pyvr_b200.cuda_renderer import VolumeRenderer

from gl_scenes import scenes
from scenes import assert_parity, c1_scene, turntable_camera, viridis_lut

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ogl.available(), reason="Mesa software libGL (Nsight Compute tree) not on this box")]

ALL = {k: v for k, v in scenes().items() if k != "surface_coin_flip"}


def _cuda(vol, cam, light, cfg, lut, w, h, **kw):
    with VolumeRenderer(w, h, config=cfg, light=light, **kw) as r:
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_lut(lut)
        return np.frombuffer(r.render(), dtype=np.uint8).reshape(h, w, 4).copy()


@pytest.mark.parametrize("name", sorted(ALL))
@pytest.mark.parametrize("mode", ["fast", "strict"])
def test_cuda_matches_the_reference_shader(name, mode):
    vol, cam, light, cfg, lut, w, h = ALL[name]
    want = ogl.render(vol, cam, light, cfg, lut, w, h)
    got = _cuda(vol, cam, light, cfg, lut, w, h, strict=(mode == "strict"))
    m = assert_parity(got, want)
    assert m["max_abs"] <= 2, m
    assert want[..., 3].max() > 20


def test_c1_full_size_against_the_shader():
    """BASELINE config C1 at its own size: 128^3 double_sphere + normals, 512x512, balanced, iso view."""
    data = create_sample_volume(128, "double_sphere")
    vol, light, lut = c1_scene(128, normals=oracle.normals(data))
    cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
    want = ogl.render(vol, cam, light, cfg, lut, 512, 512)
    for kw in (dict(), dict(texel_format="f16"), dict(hardware_filtering=True)):
        got = _cuda(vol, cam, light, cfg, lut, 512, 512, **kw)
        m = assert_parity(got, want)
        assert m["max_abs"] <= (2 if not kw else 3), (kw, m)


def test_c3_view_against_the_shader():
    """The headline scene (bounds +-1, high_quality, turntable view, viridis + linear(0, 0.1)) on a 256^3 volume at
    960x540 -- a quarter of C3's pixels and an eighth of its voxels keep the llvmpipe frame to a few seconds."""
    data = create_sample_volume(256, "double_sphere")
    vol = Volume(data=data, normals=oracle.normals(data), min_bounds=np.array([-1, -1, -1], np.float32),
                 max_bounds=np.array([1, 1, 1], np.float32))
    light, cfg, lut = Light.directional([1, -1, 0]), RenderConfig.high_quality(), viridis_lut(0.0, 0.1)
    cam = turntable_camera(40)
    want = ogl.render(vol, cam, light, cfg, lut, 960, 540)
    got = _cuda(vol, cam, light, cfg, lut, 960, 540)
    m = assert_parity(got, want)
    assert m["max_abs"] <= 2, m
