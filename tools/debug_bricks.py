import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from pyvr_b200 import Camera, RenderConfig, create_sample_volume
from pyvr_b200 import multi_gpu as mg
from pyvr_b200.cuda_renderer import VolumeRenderer
from scenes import c1_scene
W,H=320,240
data = create_sample_volume(128, "double_sphere")
vol, light, lut = c1_scene(128, normals=oracle.normals(data))
cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
def run(load, ess=True):
    with VolumeRenderer(W,H,config=cfg,light=light, empty_space_skipping=ess) as r:
        load(r); r.set_camera(cam); r.set_lut(lut)
        a = r.render_accum(); return r.stats, a
s,a0 = run(lambda r: r.load_volume(vol)); print('whole', s)
s,a1 = run(lambda r: r.load_brick(vol.data, vol.normals, data.shape, (0,0,0),(0,0,0),data.shape, vol.min_bounds, vol.max_bounds)); print('one brick', s, np.abs(a1-a0).max())
for world in (2,):
  tot=0
  for rank in range(world):
    b = mg.brick_of_rank(data.shape, rank, world)
    for ess in (True, False):
        s,a = run(lambda r: r.load_brick(vol.data[b.slices()], vol.normals[b.slices()], data.shape, b.origin, b.own_lo, b.own_hi, vol.min_bounds, vol.max_bounds), ess)
        print(world, rank, b, 'ess',ess, s)
