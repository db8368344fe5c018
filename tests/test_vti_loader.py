"""The std-lib .vti reader and ``load_vtk_volume`` (reference
pyvr/dataloaders/vtk_loader.py:15-146, tests/test_dataloaders/test_vtk_loader.py:14-76)."""

import hashlib
import json
import os

import numpy as np
import pytest

from pyvr_b200.dataloaders import load_vtk_volume, read_vti

from vti_writer import write_vti

REF_DATA = "/root/reference/example_data"
VARIANTS = [
    dict(fmt="appended", encoding="base64", compressed=True, header_type="UInt32"),   # the reference fixtures' form
    dict(fmt="appended", encoding="base64", compressed=False, header_type="UInt32"),
    dict(fmt="appended", encoding="base64", compressed=True, header_type="UInt64"),
    dict(fmt="appended", encoding="raw", compressed=True, header_type="UInt32"),
    dict(fmt="appended", encoding="raw", compressed=False, header_type="UInt64"),
    dict(fmt="binary", compressed=True, header_type="UInt32"),
    dict(fmt="binary", compressed=False, header_type="UInt32"),
    dict(fmt="ascii", compressed=False),
]


@pytest.fixture(scope="module")
def vti_golden(golden_dir):
    vols = np.load(os.path.join(golden_dir, "vti_volumes.npz"))
    meta = json.load(open(os.path.join(golden_dir, "vti_meta.json")))
    return vols, meta


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: "-".join(str(x) for x in v.values()))
def test_reader_round_trips_every_container_variant(tmp_path, variant):
    rng = np.random.default_rng(7)
    dims = (9, 6, 5)   # ragged on purpose; > 32 KiB blocks are covered by the fixture-sized test below
    a = rng.random(dims[0] * dims[1] * dims[2]).astype(np.float32)
    b = rng.integers(0, 1000, dims[0] * dims[1] * dims[2]).astype(np.int16)
    path = write_vti(tmp_path / "v.vti", {"Scalars_": a, "other": b}, dims, spacing=(1.0, 2.0, 0.5), **variant)
    got_dims, spacing, arrays = read_vti(path)
    assert got_dims == dims and spacing == (1.0, 2.0, 0.5)
    assert np.array_equal(arrays["Scalars_"][0], a) and arrays["Scalars_"][1] == 1
    assert np.array_equal(arrays["other"][0], b)


def test_multi_block_stream_matches_the_fixture_payload(tmp_path, vti_golden):
    """fuel.vti is 32 zlib blocks of 32 KiB: rebuild it in the fixture's own container form."""
    vols, meta = vti_golden
    flat = vols["fuel"].astype(np.float32).ravel()
    path = write_vti(tmp_path / "fuel.vti", {"Scalars_": flat}, meta["fuel"]["dims_xyz"])
    _, _, arrays = read_vti(path)
    assert hashlib.sha256(arrays["Scalars_"][0].astype("<f4").tobytes()).hexdigest() == meta["fuel"]["sha256_f32"]


def test_load_vtk_volume_semantics(tmp_path, vti_golden):
    """What the reference's loader tests pin: shape (nz,ny,nx), float32, [0,1] range, +-1 bounds, name."""
    vols, meta = vti_golden
    for name in ("fuel", "hydrogen"):
        flat = vols[name].astype(np.float32).ravel()
        path = write_vti(tmp_path / f"{name}.vti", {"Scalars_": flat}, meta[name]["dims_xyz"])
        vol = load_vtk_volume(path, compute_normals=False)
        n = meta[name]["dims_xyz"][0]
        assert vol.data.shape == (n, n, n) and vol.data.dtype == np.float32
        assert vol.data.min() == 0.0 and vol.data.max() == 1.0
        assert np.array_equal(vol.data, (vols[name].astype(np.float32) / np.float32(meta[name]["max"])))
        assert np.allclose(vol.min_bounds, -1.0) and np.allclose(vol.max_bounds, 1.0)
        assert vol.name == f"{name}.vti(Scalars_)" and not vol.has_normals
        assert int(np.count_nonzero(vol.data)) == meta[name]["nonzero"]


def test_bounds_follow_physical_extent_and_constant_volume(tmp_path):
    dims = (8, 4, 2)
    path = write_vti(tmp_path / "a.vti", {"Scalars_": np.arange(64, dtype=np.float32)}, dims, spacing=(1.0, 1.0, 4.0))
    vol = load_vtk_volume(path, compute_normals=False)
    assert vol.data.shape == (2, 4, 8)                       # (nz, ny, nx)
    # physical = (8, 4, 8) -> longest side spans [-1, 1]; bounds stay in VTK (x, y, z) order
    assert np.allclose(vol.max_bounds, [1.0, 0.5, 1.0]) and np.allclose(vol.min_bounds, [-1.0, -0.5, -1.0])
    path = write_vti(tmp_path / "c.vti", {"Scalars_": np.full(64, 3.0, np.float32)}, dims)
    assert not load_vtk_volume(path, compute_normals=False).data.any()   # constant volume -> zeros


def test_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        load_vtk_volume(tmp_path / "missing.vti")
    path = write_vti(tmp_path / "a.vti", {"density": np.zeros(8, np.float32)}, (2, 2, 2))
    with pytest.raises(ValueError, match="Scalar array 'Scalars_' not found"):
        load_vtk_volume(path, compute_normals=False)
    assert load_vtk_volume(path, scalars_name="density", compute_normals=False).data.shape == (2, 2, 2)
    path = write_vti(tmp_path / "m.vti", {"Scalars_": np.zeros(24, np.float32)}, (2, 2, 2), components={"Scalars_": 3})
    with pytest.raises(ValueError, match="Multi-component scalars not supported"):
        load_vtk_volume(path, compute_normals=False)
    (tmp_path / "bad.vti").write_text("<VTKFile type='PolyData'></VTKFile>")
    with pytest.raises(ValueError):
        load_vtk_volume(tmp_path / "bad.vti", compute_normals=False)


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference fixtures only exist in the build container")
def test_reference_fixtures_decode_to_the_golden_arrays(vti_golden):
    vols, meta = vti_golden
    for name in ("fuel", "hydrogen"):
        dims, spacing, arrays = read_vti(os.path.join(REF_DATA, f"{name}.vti"))
        flat = arrays["Scalars_"][0]
        assert list(dims) == meta[name]["dims_xyz"] and list(spacing) == meta[name]["spacing_xyz"]
        assert hashlib.sha256(flat.astype("<f4").tobytes()).hexdigest() == meta[name]["sha256_f32"]
        assert np.array_equal(flat.reshape(vols[name].shape), vols[name].astype(np.float32))
