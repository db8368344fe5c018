"""K2 (compute_normal_volume on the GPU) vs the reference's numpy output (golden) and the oracle.

Bar: abs(delta) <= 1e-5 * max(1, abs(ref)) per component (BASELINE.json); in practice bit-exact,
which is asserted too because the kernel performs the same binary32 operations in the same order."""

import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from pyvr_b200 import Volume, compute_normal_volume, create_sample_volume
from pyvr_b200.cuda_renderer import _cabi

pytestmark = pytest.mark.gpu


def _check(got, want, name):
    assert got.dtype == np.float32 and got.shape == want.shape, name
    tol = 1e-5 * np.maximum(1.0, np.abs(want))
    assert np.all(np.abs(got - want) <= tol), name
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{name}: not bit-exact"


def test_golden_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "normals.npz"))
    for key in [k[:-4] for k in z.files if k.endswith("__in")]:
        _check(compute_normal_volume(z[key + "__in"]), z[key + "__out"], key)


def test_double_sphere_128_hash(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    got = compute_normal_volume(create_sample_volume(128, "double_sphere"))
    assert hashlib.sha256(got.tobytes()).hexdigest() == meta["normal_volume_sha256"]["double_sphere_128"]


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 5, 8), (7, 1, 4), (3, 3, 3), (5, 6, 7), (16, 12, 20),
                                   (33, 17, 64), (64, 64, 64), (2, 2, 4),
                                   # TMA path (n2 % 4 == 0): partial tiles in y and z, chunk seams (32 planes), tiny volumes
                                   (2, 2, 8), (3, 9, 8), (35, 10, 12), (65, 19, 140), (34, 8, 256), (70, 9, 260), (97, 33, 132)])
def test_ragged_and_vector_paths_vs_oracle(shape):
    rng = np.random.default_rng(sum(shape))
    vol = (rng.standard_normal(shape) * 3).astype(np.float32)
    _check(compute_normal_volume(vol), oracle.normals(vol), str(shape))


@pytest.mark.parametrize("scale", [1e-45, 1e-38, 1e-30, 1e-26, 1e-20, 1e-16, 1e-12, 1e-8, 1.0, 1e6, 1e10, 1e19, 1e30])
def test_extreme_magnitudes_stay_bit_exact(scale):
    """The TMA kernel computes sqrt and the three quotients with nvcc's own fast-path sequences, branch-free, and
    falls back to sqrtf() and "/" outside the range on which they are exact (normals.cu, finish_fast).  Sweep the
    gradient magnitude across that range's edges -- denormals, 2^-100, 2^-75, huge values -- with mixed magnitudes
    inside one voxel, exact zeros, negative zeros and flat runs."""
    rng = np.random.default_rng(int(-np.log10(scale) + 50))
    shape = (12, 16, 64)
    base = rng.standard_normal(shape).astype(np.float32)
    mix = np.float32(10.0) ** rng.integers(-12, 1, shape).astype(np.float32)       # up to 12 decades inside a stencil
    data = (base * mix * np.float32(scale)).astype(np.float32)
    data[2:5, 3:9, 8:40] = np.float32(0.0)                                         # flat: zero gradients
    data[6:8, :, 16:24] = np.float32(-0.0)
    data[9, 4:12, :] = data[9, 4, 0]                                               # flat along one axis only
    with np.errstate(all="ignore"):
        want = oracle.normals(data)
        got = compute_normal_volume(data)
    finite = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got.view(np.uint32)[finite], want.view(np.uint32)[finite]), f"scale {scale}: not bit-exact"


def test_relaxed_opt_in_stays_within_the_tolerance():
    """compute_normal_volume(relaxed=True): g * (1/norm) instead of three correctly rounded quotients."""
    data = create_sample_volume(96, "double_sphere")
    want = oracle.normals(data)
    got = compute_normal_volume(data, relaxed=True)
    assert np.all(np.abs(got - want) <= 1e-5 * np.maximum(1.0, np.abs(want)))
    ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2 and np.array_equal(np.signbit(got), np.signbit(want))


def test_non_finite_voxels_propagate_like_numpy():
    data = np.random.default_rng(3).random((8, 8, 32)).astype(np.float32)
    data[3, 3, 7] = np.inf
    data[5, 2, 20] = np.nan
    data[1, 6, 12] = np.float32(3e38)
    with np.errstate(all="ignore"):
        want = oracle.normals(data)
        got = compute_normal_volume(data)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got.view(np.uint32)[ok], want.view(np.uint32)[ok])


def test_large_volume_vs_oracle_and_timing():
    vol = create_sample_volume(256, "helix")
    got, ms = _cabi.compute_normals_host(vol, return_ms=True)
    _check(got, oracle.normals(vol), "helix_256")
    assert 0 < ms < 50


@pytest.mark.parametrize("shape", [(37, 300, 384), (70, 250, 251), (16, 512, 512), (17, 512, 512), (1, 2048, 2048)])
def test_host_pipeline_slabs_are_bit_exact(shape):
    """Host arrays of >= 4 M voxels travel in 16-plane slabs (+ halo planes) through pinned staging
    (abi.cu, normals_host_pipeline): every plane must come out as in the whole-volume call -- slab counts that do
    and do not divide the volume, a ragged fastest axis (scalar kernel), a single plane."""
    rng = np.random.default_rng(sum(shape))
    vol = rng.random(shape, dtype=np.float32)
    vol[:, :7, :9] = 0.0
    got, ms = _cabi.compute_normals_host(vol, return_ms=True)
    _check(got, oracle.normals(vol), f"pipeline {shape}")
    assert ms > 0


def test_volume_compute_normals_method_and_dtype_cast():
    v = Volume(data=create_sample_volume(24, "torus").astype(np.float64))
    v.compute_normals()
    assert v.normals.shape == (24, 24, 24, 3) and v.normals.dtype == np.float32
    _check(v.normals, oracle.normals(v.data.astype(np.float32)), "torus_24")
    with pytest.raises(ValueError, match="Unsupported method"):
        v.compute_normals("sobel")
    with pytest.raises(ValueError, match="must be 3D"):
        compute_normal_volume(np.zeros((4, 4), np.float32))


def test_brick_and_slab_normals_match_the_whole_volume():
    """Multi-GPU split of K2 (SURVEY.md section 8 e): a brick / slab plus a one-voxel halo gives exactly the slice
    of the whole normal volume (here every 'rank' runs on the one visible GPU)."""
    from pyvr_b200 import multi_gpu as mg

    vol = create_sample_volume(64, "double_sphere")
    want = compute_normal_volume(vol)
    for b in mg.split_bricks(vol.shape, mg.brick_grid(8)):
        got = mg.brick_normals(vol, b)
        assert np.array_equal(got.view(np.uint32), want[b.slices()].view(np.uint32)), b.coord
    for rank in range(3):
        x0, x1 = mg.slab_range(64, rank, 3)
        block, inner = mg.halo_block(vol, (x0, 0, 0), (x1, 64, 64))
        got = compute_normal_volume(block)[inner]
        assert np.array_equal(got.view(np.uint32), want[x0:x1].view(np.uint32)), rank
