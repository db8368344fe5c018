#!/usr/bin/env python
"""The reference's own benchmark protocol (examples/benchmark.py:28-30,55-64,89-110) on this backend.

256^3 double_sphere + normals, bounds +-1, 512x512, camera az 45 / el 30 / distance 3, Light.default(),
plasma colour TF, linear(0, 0.1) opacity TF; per preset: a fresh renderer, 1 warm-up render(), then 10 timed
render() calls with time.perf_counter() -- each call returns the RGBA8 frame as host bytes, exactly like the
reference (read-back included).  Prints one JSON line; BASELINE.md section 1 holds the numbers the
reference's authors published for the same protocol on their (unnamed) GPU.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig, Volume,
                       compute_normal_volume, create_sample_volume)
from pyvr_b200.cuda_renderer import VolumeRenderer

PUBLISHED_MS = {"preview": "2-3", "fast": "3-4", "balanced": "4-5", "high_quality": "6-7", "ultra_quality": "24-25"}

data = create_sample_volume(256, "double_sphere")
t0 = time.perf_counter()
normals = compute_normal_volume(data)
normals_ms = (time.perf_counter() - t0) * 1e3
volume = Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32), max_bounds=np.array([1, 1, 1], np.float32))
camera = Camera.from_spherical(target=np.array([0.0, 0.0, 0.0]), distance=3.0, azimuth=np.pi / 4, elevation=np.pi / 6, roll=0.0)
light = Light.default()
ctf, otf = ColorTransferFunction.from_colormap("plasma"), OpacityTransferFunction.linear(0.0, 0.1)
rows = {}
for name in ("preview", "fast", "balanced", "high_quality", "ultra_quality"):
    r = VolumeRenderer(512, 512, config=getattr(RenderConfig, name)(), light=light)
    r.load_volume(volume)
    r.set_camera(camera)
    r.set_transfer_functions(ctf, otf)
    r.render()
    times = []
    for _ in range(10):
        t = time.perf_counter()
        frame = r.render()
        times.append((time.perf_counter() - t) * 1e3)
    st = r.stats
    rows[name] = {"mean_ms": float(np.mean(times)), "std_ms": float(np.std(times)), "fps": 1000 / float(np.mean(times)),
                  "kernel_ms": st["kernel_ms"], "samples": st["samples"], "published_ms_v0.3.4": PUBLISHED_MS[name],
                  "frame_bytes": len(frame)}
    r.close()
print(json.dumps({"protocol": "examples/benchmark.py (256^3, 512x512, 1 warm-up + 10 timed render() incl. read-back)",
                  "compute_normal_volume_256_ms_incl_copies": normals_ms, "presets": rows}))
