"""Oracle vs the reference shader on Mesa llvmpipe, per scene of tests/gl_scenes.py -> profiles/r02_gl_pin.json.

    python tools/gl_pin_report.py [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle   # noqa: E402
import oracle.gl as ogl   # noqa: E402
from gl_scenes import scenes   # noqa: E402
from scenes import image_metrics   # noqa: E402

rows = {}
for name, (vol, cam, light, cfg, lut, w, h) in sorted(scenes().items()):
    t0 = time.perf_counter()
    got = ogl.render(vol, cam, light, cfg, lut, w, h)
    t1 = time.perf_counter()
    want, _, st = oracle.render(vol, cam, light, cfg, lut, w, h)
    m = image_metrics(want, got)
    rows[name] = {**m, "samples": st["samples"], "rays_hit": st["rays_hit"], "size": [w, h], "gl_seconds": round(t1 - t0, 3)}
    print(f"{name:22s} max|d|={m['max_abs']}  within2={m['frac_within_2']:.5f}  identical={m['frac_identical']:.4f}  psnr={m['psnr_db']:.1f} dB")
r = ogl.GLReference(8, 8)
out = {"what": "oracle/pyvr_oracle.c vs pyvr/shaders/volume.frag.glsl (verbatim) executed by Mesa llvmpipe through oracle/gl "
               "(ctypes replay of pyvr/moderngl_renderer/manager.py); RGBA8 frames, per-channel |delta| in 1/255 units",
       "gl": r.info, "scenes": rows}
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_gl_pin.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
