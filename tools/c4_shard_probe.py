"""One GPU: how much march time does image-space sharding itself cost on C4?  (tools only)

Marches the tile groups of single ranks of an 8-way deal on ONE device and compares the kernel times with the
full-frame march: sum over ranks / full frame = what the deal loses to smaller coherent regions (halo lines fetched
by several ranks, shorter waves), before any exchange.  Usage: python tools/c4_shard_probe.py [size] [ranks]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from pyvr_b200 import Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig, build_rgba_lut
from pyvr_b200.cuda_renderer import VolumeRenderer

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W, H = 3840, 2160
r = VolumeRenderer(W, H, config=RenderConfig.ultra_quality(), light=Light.directional([1, -1, 0]), texel_format="f16")
stream = torch.cuda.Stream()
r.set_stream(stream.cuda_stream)
r.generate_volume(size, "double_sphere", (-0.5,) * 3, (0.5,) * 3)
r.set_lut(build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.1)))
r.set_camera(Camera.isometric_view(distance=3.0))
frame = torch.zeros(W * H * 4, dtype=torch.uint8, device="cuda")


def march_ms(reps=4):
    best = 1e9
    for _ in range(reps):
        r.render_to_device(frame.data_ptr())
        best = min(best, r.stats["kernel_ms"])
    return best


full = march_ms()
print(f"full frame: {full:.3f} ms")
for shift in (0, 1, 2, 3, 4):
    times = []
    for rank in range(world):
        r.set_pixel_shard(rank, world, in_place=True, group_shift=shift)
        times.append(march_ms(3))
    r.set_pixel_shard(0, 1)
    t = np.array(times)
    print(f"group_shift {shift} ({16 << shift}x{8 << shift} px groups): per rank mean {t.mean():.3f} max {t.max():.3f} ms, "
          f"sum/full {t.sum() / full:.2f}, speed-up of the slowest rank {full / t.max():.2f}x")
