"""Renders tests/gl_scenes.py (small subset) with the reference's GLSL on Mesa llvmpipe (oracle/gl) and stores
the RGBA8 frames in tests/golden/gl_frames.npz.  Run in the build container:

    python tests/golden/make_gl_golden.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle.gl as ogl   # noqa: E402
from gl_scenes import scenes   # noqa: E402

frames = {}
for name, (vol, cam, light, cfg, lut, w, h) in scenes(small=True).items():
    frames[name] = ogl.render(vol, cam, light, cfg, lut, w, h)
    print(name, frames[name].shape, int(frames[name][..., 3].max()))
np.savez_compressed(os.path.join(HERE, "gl_frames.npz"), **frames)
r = ogl.GLReference(8, 8)
print(r.info)
