"""world_size 2 and 4 `gloo` runs of the distributed executors (CPU tensors, numpy-style `over`): the same
SortLastSession / reduce_tile_frames code paths the NCCL runs use."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyvr_b200 import multi_gpu as mg

from test_multi_gpu_plan import geometric_partials


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torch_over(front, back, term=2.0):
    out = front + (1.0 - front[:, 3:4]) * back
    hide = front[:, 3] >= term
    out[hide] = front[hide]
    return out


def _worker(rank, world, port, n, shape, cam, result_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        partial = geometric_partials(shape, world, cam, n, seed=7)[0][rank]
        session = mg.SortLastSession(shape, [-1, -1, -1], [1, 1, 1], n, exchange="nccl", over=torch_over)
        cam_world = (np.asarray(cam) + 0.5) / np.asarray(shape) * 2.0 - 1.0     # inverse of camera_in_voxels
        (lo, hi), piece = session.composite(cam_world, image=torch.from_numpy(partial.copy()))
        np.save(os.path.join(result_dir, f"piece_{rank}.npy"), piece.numpy())
        np.save(os.path.join(result_dir, f"range_{rank}.npy"), np.array([lo, hi]))
        # view sharding + tile-frame reduction on the same group
        frame = torch.zeros(64, dtype=torch.uint8)
        frame[rank::world] = rank + 1
        mg.reduce_tile_frames(frame, dst=0)
        if rank == 0:
            np.save(os.path.join(result_dir, "frame.npy"), frame.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cam", [(2, (-100.0, 20.0, 30.0)), (2, (400.0, 20.0, 30.0)), (4, (30.0, 900.0, 64.0))])
def test_sort_last_session_over_gloo(tmp_path, world, cam):
    n, shape = 5000, (128, 128, 128)
    mp.spawn(_worker, args=(world, _free_port(), n, shape, cam, str(tmp_path)), nprocs=world, join=True)
    got = np.zeros((n, 4), np.float32)
    covered = np.zeros(n, np.int32)
    for r in range(world):
        lo, hi = np.load(tmp_path / f"range_{r}.npy")
        got[lo:hi] = np.load(tmp_path / f"piece_{r}.npy")
        covered[lo:hi] += 1
    assert (covered == 1).all()
    want = geometric_partials(shape, world, cam, n, seed=7)[1]
    assert want[:, 3].max() > 0.3 and np.allclose(got, want, atol=3e-6)
    frame = np.load(tmp_path / "frame.npy")
    assert np.array_equal(frame, np.tile(np.arange(1, world + 1, dtype=np.uint8), 64 // world + 1)[:64])


def _normals_worker(rank, world, port, shape, result_dir):
    import oracle

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = (np.random.default_rng(5).standard_normal(shape) * 2).astype(np.float32)
        got = mg.compute_normal_volume_sharded(vol, normals_fn=oracle.normals)
        np.save(os.path.join(result_dir, f"normals_{rank}.npy"), got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (9, 6, 8)), (4, (10, 5, 12)), (2, (2, 3, 4))])
def test_sharded_normal_volume_is_bit_identical(tmp_path, world, shape):
    """Slab split of compute_normal_volume with a one-plane halo (SURVEY.md section 8 e): every rank ends up with
    exactly the single-process normal volume."""
    import oracle

    mp.spawn(_normals_worker, args=(world, _free_port(), shape, str(tmp_path)), nprocs=world, join=True)
    vol = (np.random.default_rng(5).standard_normal(shape) * 2).astype(np.float32)
    want = oracle.normals(vol)
    for r in range(world):
        got = np.load(tmp_path / f"normals_{r}.npy")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), r


def test_brick_normals_equal_the_sliced_normal_volume():
    import oracle

    vol = (np.random.default_rng(9).standard_normal((12, 10, 16)) * 2).astype(np.float32)
    want = oracle.normals(vol)
    for world in (2, 4, 8):
        for b in mg.split_bricks(vol.shape, mg.brick_grid(world)):
            got = mg.brick_normals(vol, b, normals_fn=oracle.normals)
            assert np.array_equal(got.view(np.uint32), want[b.slices()].view(np.uint32)), (world, b.coord)
