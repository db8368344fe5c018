"""Multi-GPU paths on the device (SURVEY.md section 8 e): image-tile shards and sort-last bricks.

The 1-GPU tests emulate P ranks inside one process -- the very schedule (`binary_swap_plan`) and kernels
(`composite_over`, `finalize_rgba8`, the BRICK march) the distributed runs use -- and hold the result to the
un-partitioned render and to the oracle.  The tests at the bottom spawn one process per GPU (NCCL and the
CUDA-IPC peer path) and are skipped on a single-GPU box.
"""

import ctypes
import dataclasses
import os
import socket

import numpy as np
import pytest

import oracle
from pyvr_b200 import Camera, Light, RenderConfig, Volume, create_sample_volume
from pyvr_b200 import multi_gpu as mg
from pyvr_b200.cuda_renderer import VolumeRenderer, _cabi

from scenes import assert_parity, c1_scene, image_metrics, viridis_lut

pytestmark = pytest.mark.gpu

W, H = 320, 240


class DevImage:
    """A float4 image in device memory that slices by pixel (what composite_in_process indexes)."""

    def __init__(self, buf, ptr, n):
        self.buf, self.ptr, self.n = buf, ptr, n

    def __getitem__(self, s):
        lo, hi, _ = s.indices(self.n)
        return DevImage(self.buf, self.ptr + lo * 16, hi - lo)


def dev_over(front, back, term=0.99):
    out = _cabi.DeviceBuffer(front.n * 16)
    _cabi.composite_over(0, front.ptr, back.ptr, out.ptr, front.n, term)
    _cabi.stream_synchronize(0)
    return DevImage(out, out.ptr, front.n)


@pytest.fixture(scope="module")
def scene():
    data = create_sample_volume(128, "double_sphere")
    vol, light, lut = c1_scene(128, normals=oracle.normals(data))
    return vol, light, lut


def _whole(vol, light, lut, cam, cfg):
    with VolumeRenderer(W, H, config=cfg, light=light) as r:
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_lut(lut)
        frame = np.frombuffer(r.render(), np.uint8).reshape(H, W, 4).copy()
        stats = r.stats
        accum = r.render_accum()
    return frame, accum, stats


def _in_box_samples(vol, light, lut, cam, cfg):
    """Samples of all rays with the stop rule off: what the bricks execute together when no brick saturates
    on its own (a brick cannot see the alpha accumulated in the bricks in front of it)."""
    never = dataclasses.replace(cfg, early_ray_termination=False)
    with VolumeRenderer(W, H, config=never, light=light, honor_config_termination=True) as r:
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_lut(lut)
        r.render()
        return r.stats["samples"]


def _bricked(vol, light, lut, cam, cfg, world, texel_format="f32"):
    """Render every brick into a device partial image; returns (partials, per-brick stats)."""
    shape = vol.data.shape
    partials, stats = [], []
    for rank in range(world):
        b = mg.brick_of_rank(shape, rank, world)
        with VolumeRenderer(W, H, config=cfg, light=light, texel_format=texel_format) as r:
            r.load_brick(vol.data[b.slices()], vol.normals[b.slices()] if vol.has_normals else None,
                         shape, b.origin, b.own_lo, b.own_hi, vol.min_bounds, vol.max_bounds)
            r.set_camera(cam)
            r.set_lut(lut)
            buf = _cabi.DeviceBuffer(W * H * 16)
            r.render_accum_to_device(buf.ptr)
            stats.append(r.stats)
        partials.append(DevImage(buf, buf.ptr, W * H))
    return partials, stats


def _composite(partials, vol, cam, term=0.99):
    position, _ = cam.get_camera_vectors()
    cam_vox = mg.camera_in_voxels(position, vol.min_bounds, vol.max_bounds, vol.data.shape)
    pieces = mg.composite_in_process(partials, vol.data.shape, cam_vox, lambda f, b: dev_over(f, b, term), W * H)
    accum = np.zeros((W * H, 4), np.float32)
    frame = np.zeros((W * H, 4), np.uint8)
    for (lo, hi), img in pieces:
        accum[lo:hi] = img.buf.to_host(np.float32, img.ptr - img.buf.ptr, (hi - lo) * 16).reshape(-1, 4)
        out = _cabi.DeviceBuffer((hi - lo) * 4)
        _cabi.finalize_rgba8(0, img.ptr, out.ptr, hi - lo)
        _cabi.stream_synchronize(0)
        frame[lo:hi] = out.to_host(np.uint8).reshape(-1, 4)
    return frame.reshape(H, W, 4), accum.reshape(H, W, 4)


CAMS = {
    "iso": lambda: Camera.isometric_view(distance=3.0),
    "front": lambda: Camera.front_view(distance=3.0),             # looks along an axis: rays parallel to split planes
    "rolled": lambda: Camera(azimuth=1.1, elevation=-0.4, roll=0.7, distance=2.2),
    "inside": lambda: Camera(azimuth=0.3, elevation=0.2, distance=0.3),
    "behind": lambda: Camera(azimuth=3.9, elevation=0.5, distance=2.5),
}


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("cam_name", list(CAMS))
def test_bricks_composite_to_the_whole_volume(scene, world, cam_name):
    vol, light, lut = scene
    cam, cfg = CAMS[cam_name](), RenderConfig.balanced()
    want, want_acc, want_stats = _whole(vol, light, lut, cam, cfg)
    partials, stats = _bricked(vol, light, lut, cam, cfg, world)
    # every sample of every ray is owned by exactly one brick: with no brick saturating on its own the bricks
    # execute every in-box sample between them -- including the few the single-pass stop rule cuts
    total, in_box = sum(s["samples"] for s in stats), _in_box_samples(vol, light, lut, cam, cfg)
    assert want_stats["samples"] <= total <= in_box
    if all(s["rays_terminated"] == 0 for s in stats):
        assert total == in_box
    got, got_acc = _composite(partials, vol, cam)
    # same samples, same arithmetic per sample; only the association order of the sums differs.  Rays
    # that saturate are cut at alpha = 0.99 by the merge (composite.cu) where the single pass stops within
    # one sample's contribution above it.
    sat = want_acc[..., 3] >= 0.985
    assert np.abs(got_acc - want_acc)[~sat].max() < 2e-5
    assert np.abs(got_acc - want_acc).max() < 0.006
    m = image_metrics(got, want)
    assert m["max_abs"] <= 2 and m["frac_within_1"] >= 0.999, m
    oracle_img, _, _ = oracle.render(vol, cam, light, cfg, lut, W, H)
    assert_parity(got, oracle_img)


def test_bricks_with_saturating_rays_report_error(scene):
    """Opaque transfer function (single samples with alpha up to 1): most rays stop early, many of them inside
    a back brick, and the single pass can overshoot 0.99 by a whole sample.  The clipped merge of
    composite.cu keeps the frame within 3/255 (99.5 % of pixels within 2); the relay mode is exact."""
    vol, light, _ = scene
    lut = viridis_lut(0.0, 1.0)
    lut[:, 3] = np.minimum(lut[:, 3] * 6.0, 1.0)
    cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
    want, want_acc, want_stats = _whole(vol, light, lut, cam, cfg)
    assert want_stats["rays_terminated"] > 1000
    partials, stats = _bricked(vol, light, lut, cam, cfg, 8)
    got, _ = _composite(partials, vol, cam)
    m = image_metrics(got, want)
    assert m["max_abs"] <= 3 and m["frac_within_2"] >= 0.995, m
    unsaturated = want_acc[..., 3] < 0.97
    assert np.abs(got.astype(int) - want.astype(int))[unsaturated].max() <= 1


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("cam_name", ["iso", "rolled", "inside"])
@pytest.mark.parametrize("opaque", [False, True])
def test_relay_is_bit_identical_to_the_single_gpu_march(scene, world, cam_name, opaque):
    """Passing one accumulating image through the bricks in visibility order repeats the single-GPU march
    operation for operation: float accumulators, RGBA8 bytes and sample counts are all EQUAL, stop rule
    included (the opaque variant terminates most rays, many inside a back brick)."""
    vol, light, lut = scene
    if opaque:
        lut = viridis_lut(0.0, 1.0)
        lut[:, 3] = np.minimum(lut[:, 3] * 6.0, 1.0)
    cam, cfg = CAMS[cam_name](), RenderConfig.balanced()
    want, want_acc, want_stats = _whole(vol, light, lut, cam, cfg)
    position, _ = cam.get_camera_vectors()
    cam_vox = mg.camera_in_voxels(position, vol.min_bounds, vol.max_bounds, vol.data.shape)
    order = mg.relay_order(world, vol.data.shape, cam_vox)
    assert sorted(order) == list(range(world))
    image = _cabi.DeviceBuffer(W * H * 16)
    samples = terminated = 0
    for pos, rank in enumerate(order):
        b = mg.brick_of_rank(vol.data.shape, rank, world)
        with VolumeRenderer(W, H, config=cfg, light=light) as r:
            r.load_brick(vol.data[b.slices()], vol.normals[b.slices()], vol.data.shape, b.origin, b.own_lo, b.own_hi,
                         vol.min_bounds, vol.max_bounds)
            r.set_camera(cam)
            r.set_lut(lut)
            r.render_accum_relay(image.ptr if pos else None, image.ptr)
            samples += r.stats["samples"]
            terminated += r.stats["rays_terminated"]
    got_acc = image.to_host(np.float32).reshape(H, W, 4)
    out = _cabi.DeviceBuffer(W * H * 4)
    _cabi.finalize_rgba8(0, image.ptr, out.ptr, W * H)
    _cabi.stream_synchronize(0)
    got = out.to_host(np.uint8).reshape(H, W, 4)
    assert np.array_equal(got_acc, want_acc)
    assert np.array_equal(got, want)
    assert samples == want_stats["samples"] and terminated == want_stats["rays_terminated"]
    if opaque:
        assert want_stats["rays_terminated"] > 1000


def test_half_texel_bricks_and_non_cubic_volume():
    rng = np.random.default_rng(3)
    data = rng.random((48, 40, 72)).astype(np.float32) * 0.4
    data[10:30, 8:30, 20:60] += 0.5
    vol = Volume(data=data, normals=oracle.normals(data), min_bounds=np.array([-0.6, -0.5, -0.9], np.float32),
                 max_bounds=np.array([0.6, 0.5, 0.9], np.float32))
    light, lut = Light.directional([1, -1, 0]), viridis_lut(0.0, 0.03)
    cam, cfg = Camera(azimuth=0.7, elevation=0.3, distance=3.0), RenderConfig.balanced()
    # a whole volume uploaded through the brick entry point (one brick = everything) is the sane
    # data[ix,iy,iz] mapping; compare the 4-brick render against it
    with VolumeRenderer(W, H, config=cfg, light=light) as r:
        r.load_brick(vol.data, vol.normals, data.shape, (0, 0, 0), (0, 0, 0), data.shape, vol.min_bounds, vol.max_bounds)
        r.set_camera(cam)
        r.set_lut(lut)
        want = np.frombuffer(r.render(), np.uint8).reshape(H, W, 4).copy()
        assert r.stats["rays_terminated"] == 0 and r.stats["rays_hit"] > 5000
        n_want = r.stats["samples"]
    assert want.any()
    for fmt in ("f32", "f16"):
        partials, stats = _bricked(vol, light, lut, cam, cfg, 4, texel_format=fmt)
        assert sum(s["samples"] for s in stats) == n_want
        got, _ = _composite(partials, vol, cam)
        if fmt == "f32":
            assert image_metrics(got, want)["max_abs"] <= 1
        else:
            assert_parity(got, want)


@pytest.mark.parametrize("count", [2, 3, 8])
def test_pixel_shards_add_up_to_the_frame_bit_for_bit(scene, count):
    vol, light, lut = scene
    cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
    w, h = 400, 300      # not a multiple of the 64-pixel tile groups
    with VolumeRenderer(w, h, config=cfg, light=light) as r:
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_lut(lut)
        want = np.frombuffer(r.render(), np.uint8).reshape(h, w, 4).copy()
        n_want = r.stats["samples"]
        total = np.zeros((h, w, 4), np.uint32)
        samples, covered = 0, np.zeros((h, w), np.int32)
        for rank in range(count):
            r.set_pixel_shard(rank, count)
            part = np.frombuffer(r.render(), np.uint8).reshape(h, w, 4)
            samples += r.stats["samples"]
            total += part
            covered += part.any(axis=-1)
        r.set_pixel_shard(0, 1)
        again = np.frombuffer(r.render(), np.uint8).reshape(h, w, 4)
    assert np.array_equal(total, want) and samples == n_want and covered.max() == 1
    assert np.array_equal(again, want)


# ---------------------------------------------------------------------------------------------------
# one process per GPU
# ---------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dist_worker(rank, world, port, out_dir, exchange):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        data = create_sample_volume(128, "double_sphere")
        vol, light, lut = c1_scene(128, normals=oracle.normals(data))
        cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
        position, _ = cam.get_camera_vectors()
        b = mg.brick_of_rank(data.shape, rank, world)
        with VolumeRenderer(W, H, config=cfg, light=light, device=rank) as r:
            # the session binds the renderer to torch's current stream (merges, flags and NCCL traffic run there)
            session = mg.SortLastSession(data.shape, vol.min_bounds, vol.max_bounds, W * H, device=rank, exchange=exchange,
                                         renderer=r)
            # the brick's normals from its own voxels + a one-voxel halo, not from the whole normal volume
            r.load_brick(vol.data[b.slices()], mg.brick_normals(vol.data, b, device=rank), data.shape, b.origin, b.own_lo,
                         b.own_hi, vol.min_bounds, vol.max_bounds)
            r.set_camera(cam)
            r.set_lut(lut)
            for it in range(3):   # several frames: buffers and peer flags are reused; the last one fuses finalize into the merge
                r.render_accum_to_device(session.image_ptr())
                piece_range, piece = session.composite(position, finalize_to=0 if (exchange == "p2p" and it == 2) else None)
                frame = session.gather_rgba8(piece_range, piece)
                if isinstance(frame, int):      # p2p: raw pointer of the IPC-shared frame buffer on rank 0
                    torch.cuda.synchronize()
                    host = np.empty(W * H * 4, np.uint8)
                    _cabi.check(_cabi.lib().pyvr_cuda_memcpy(rank, host.ctypes.data, ctypes.c_void_p(frame), W * H * 4, 2, None))
                    frame = torch.from_numpy(host)
            samples = torch.tensor([r.stats["samples"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(samples)
            # exact relay on the same bricks
            relay = mg.RelaySession(data.shape, vol.min_bounds, vol.max_bounds, W * H, device=rank)
            relay_frame = relay.render(r, position)
            if relay_frame is not None:
                np.save(os.path.join(out_dir, "relay.npy"), relay_frame.cpu().numpy().reshape(H, W, 4))
            # image tiles on the same ranks: replicated volume, reduce(SUM) of uint8 frames
            r.load_volume(vol)
            r.set_pixel_shard(rank, world)
            tiles = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
            r.render_to_device(tiles.data_ptr())
            mg.reduce_tile_frames(tiles, dst=0)
            # ... and fused: every rank's march stores its pixels straight into rank 0's frame (peer memory + flags)
            ts = mg.TileSession(r, W * H, device=rank)
            for _ in range(3):
                ptr = ts.render()
                if ptr is not None:
                    torch.cuda.synchronize()
                    fused = np.empty(W * H * 4, np.uint8)
                    _cabi.check(_cabi.lib().pyvr_cuda_memcpy(rank, fused.ctypes.data, ctypes.c_void_p(ptr), W * H * 4, 2, None))
                ts.release()
            ts.close()
        if rank == 0:
            np.save(os.path.join(out_dir, "frame.npy"), frame.cpu().numpy().reshape(H, W, 4))
            np.save(os.path.join(out_dir, "tiles.npy"), tiles.cpu().numpy().reshape(H, W, 4))
            np.save(os.path.join(out_dir, "tiles_fused.npy"), fused.reshape(H, W, 4))
            np.save(os.path.join(out_dir, "samples.npy"), samples.cpu().numpy())
        session.close()
    finally:
        dist.destroy_process_group()


def _gpu_count():
    try:
        return _cabi.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_sort_last_and_tiles_across_processes(scene, tmp_path, exchange):
    world = 8 if _gpu_count() >= 8 else 4 if _gpu_count() >= 4 else 2
    if _gpu_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_dist_worker, args=(world, _free_port(), str(tmp_path), exchange), nprocs=world, join=True)
    vol, light, lut = scene
    cam, cfg = Camera.isometric_view(distance=3.0), RenderConfig.balanced()
    want, _, want_stats = _whole(vol, light, lut, cam, cfg)
    got = np.load(tmp_path / "frame.npy")
    assert int(np.load(tmp_path / "samples.npy")[0]) == _in_box_samples(vol, light, lut, cam, cfg)
    m = image_metrics(got, want)
    assert m["max_abs"] <= 3 and m["frac_within_1"] >= 0.999, m
    assert np.array_equal(np.load(tmp_path / "tiles.npy"), want)
    assert np.array_equal(np.load(tmp_path / "tiles_fused.npy"), want)
    assert np.array_equal(np.load(tmp_path / "relay.npy"), want)
