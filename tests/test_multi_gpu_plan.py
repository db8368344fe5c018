"""Host logic of the multi-GPU partitioning (pyvr_b200/multi_gpu.py): pure-Python plans and the binary-swap
schedule, checked on CPU against a sequential `over` in an independently derived visibility order."""

import itertools

import numpy as np
import pytest

from pyvr_b200 import multi_gpu as mg


def np_over(front, back, term=0.99):
    """front over back on (n,4) premultiplied float arrays, with the brick-level stop rule of composite.cu."""
    t = (1.0 - front[:, 3:4]).astype(np.float32)
    out = front + t * back
    hide = front[:, 3] >= term
    out[hide] = front[hide]
    return out.astype(np.float32)


def geometric_partials(shape, world, cam_voxel, n_rays, seed=0):
    """Partial images from real geometry: every brick is a homogeneous box with its own colour and density;
    ray p from the camera gets, per brick, opacity 1 - exp(-sigma * chord) and premultiplied colour.  Returns
    (partials by rank, ground truth), the truth being the per-ray composite of the bricks SORTED BY ENTRY
    DISTANCE -- no plane rule, no brick-level order."""
    rng = np.random.default_rng(seed)
    grid = mg.brick_grid(world)
    cam = np.asarray(cam_voxel, np.float64)
    target = rng.uniform(0, 1, (n_rays, 3)) * np.asarray(shape)
    d = target - cam
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    bricks = {b.coord: b for b in mg.split_bricks(shape, grid)}
    partials, t_in_all = [], []
    for rank in range(world):
        b = bricks[mg.rank_to_brick(rank, world)]
        lo = np.array([(-1e9 if b.own_lo[a] == 0 else b.own_lo[a]) for a in range(3)], np.float64)
        hi = np.array([(1e9 if b.own_hi[a] == shape[a] else b.own_hi[a]) for a in range(3)], np.float64)
        lo, hi = np.maximum(lo, -0.5), np.minimum(hi, np.asarray(shape) - 0.5)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (lo - cam) / d, (hi - cam) / d
        tn = np.max(np.minimum(t0, t1), axis=1)
        tf = np.min(np.maximum(t0, t1), axis=1)
        tn = np.maximum(tn, 0.0)
        chord = np.maximum(tf - tn, 0.0)
        sigma, colour = 0.004 * (1 + rank % 3), rng.uniform(0.2, 1.0, 3)
        alpha = (1.0 - np.exp(-sigma * chord)).astype(np.float32)[:, None]
        partials.append(np.concatenate([colour.astype(np.float32)[None, :] * alpha, alpha], axis=1))
        t_in_all.append(np.where(chord > 0, tn, np.inf))
    order = np.argsort(np.stack(t_in_all, axis=1), axis=1, kind="stable")
    truth = np.zeros((n_rays, 4), np.float32)
    stacked = np.stack(partials, axis=1)                      # (rays, ranks, 4)
    for k in range(world):
        nxt = stacked[np.arange(n_rays), order[:, k]]
        truth = truth + (1.0 - truth[:, 3:4]) * nxt
    return partials, truth.astype(np.float32)


def test_view_sharding_covers_every_view_once():
    for world in (1, 2, 3, 8):
        seen = sorted(k for r in range(world) for k in mg.shard_views(360, r, world))
        assert seen == list(range(360))
    with pytest.raises(ValueError):
        mg.shard_views(10, 2, 2)


def test_brick_grid_and_rank_mapping():
    assert mg.brick_grid(1) == (1, 1, 1) and mg.brick_grid(2) == (2, 1, 1)
    assert mg.brick_grid(4) == (2, 2, 1) and mg.brick_grid(8) == (2, 2, 2) and mg.brick_grid(16) == (4, 2, 2)
    for world in (1, 2, 4, 8, 16, 64):
        grid = mg.brick_grid(world)
        coords = {mg.rank_to_brick(r, world) for r in range(world)}
        assert coords == set(itertools.product(range(grid[0]), range(grid[1]), range(grid[2])))
    with pytest.raises(ValueError):
        mg.brick_grid(6)


@pytest.mark.parametrize("shape,grid", [((128, 128, 128), (2, 2, 2)), ((64, 48, 40), (2, 2, 1)), ((9, 5, 3), (2, 1, 1)),
                                        ((512, 512, 512), (4, 2, 2))])
def test_bricks_partition_the_volume_with_upper_ghost(shape, grid):
    bricks = mg.split_bricks(shape, grid)
    assert len(bricks) == grid[0] * grid[1] * grid[2]
    owned = np.zeros(shape, dtype=np.int32)
    for b in bricks:
        owned[tuple(slice(lo, hi) for lo, hi in zip(b.own_lo, b.own_hi))] += 1
        for a in range(3):
            assert b.origin[a] == b.own_lo[a]
            # one ghost voxel above, except at the volume's outer face
            assert b.origin[a] + b.dims[a] == min(b.own_hi[a] + 1, shape[a])
    assert (owned == 1).all()


@pytest.mark.parametrize("world", [1, 2, 4, 8, 16])
def test_swap_plan_ranges_are_consistent(world):
    n = 1000 * 37 + 3          # deliberately not divisible
    pieces = []
    for rank in range(world):
        plan = mg.binary_swap_plan(rank, world, n)
        lo, hi = 0, n
        for step in plan:
            partner_step = mg.binary_swap_plan(step.partner, world, n)[step.round]
            assert partner_step.partner == rank and partner_step.keep == step.give and partner_step.give == step.keep
            assert partner_step.axis == step.axis and partner_step.plane_brick == step.plane_brick
            assert partner_step.low_side != step.low_side
            assert sorted([step.keep, step.give]) == [(lo, (lo + hi) // 2), ((lo + hi) // 2, hi)]
            lo, hi = step.keep
        pieces.append(mg.final_piece(rank, world, n))
    pieces.sort()
    assert pieces[0][0] == 0 and pieces[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))


@pytest.mark.parametrize("world", [2, 4, 8, 16])
@pytest.mark.parametrize("cam", [(-200.0, 30.0, 500.0), (20.0, 100.0, -300.0), (300.0, 300.0, 300.0), (10.0, 500.0, 70.0),
                                 (40.0, 70.0, 90.0)])
def test_binary_swap_equals_per_ray_depth_sorted_compositing(world, cam):
    shape, n = (128, 128, 128), 4096
    partials, truth = geometric_partials(shape, world, cam, n, seed=world)
    pieces = mg.composite_in_process(partials, shape, cam, lambda f, b: np_over(f, b, term=2.0), n)
    got = np.zeros((n, 4), np.float32)
    for (lo, hi), img in pieces:
        got[lo:hi] = img
    assert truth[:, 3].max() > 0.3                      # the scene is not trivially transparent
    assert np.allclose(got, truth, atol=3e-6)
    # the plane rule matters: flipping it must change the picture
    flipped = mg.composite_in_process(partials, shape, cam, lambda f, b: np_over(b, f, term=2.0), n)
    wrong = np.zeros((n, 4), np.float32)
    for (lo, hi), img in flipped:
        wrong[lo:hi] = img
    assert not np.allclose(wrong, truth, atol=1e-3)


@pytest.mark.parametrize("world", [1, 2, 4, 8, 16])
def test_relay_order_is_a_per_ray_depth_order(world):
    """Folding the partial images front to back in relay_order equals per-ray depth-sorted compositing."""
    shape, n = (128, 128, 128), 2048
    for cam in [(-200.0, 30.0, 500.0), (300.0, 300.0, 300.0), (40.0, 70.0, 90.0), (64.0, -5.0, 200.0)]:
        partials, truth = geometric_partials(shape, world, cam, n, seed=3)
        order = mg.relay_order(world, shape, cam)
        assert sorted(order) == list(range(world))
        acc = np.zeros((n, 4), np.float32)
        for rank in order:
            acc = np_over(acc, partials[rank], term=2.0)
        assert np.allclose(acc, truth, atol=3e-6)


def test_camera_in_voxels_matches_march_mapping():
    v = mg.camera_in_voxels([0.0, -1.0, 1.0], [-1, -1, -1], [1, 1, 1], (512, 512, 512))
    assert np.allclose(v, [255.5, -0.5, 511.5])
