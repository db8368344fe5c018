"""Shared scene builders for the parity tests (SURVEY.md section 8 d configs, scaled down)."""

import numpy as np

from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                       Volume, build_rgba_lut, create_sample_volume)


def viridis_lut(lo=0.0, hi=0.3, size=None):
    return build_rgba_lut(ColorTransferFunction.from_colormap("viridis"),
                          OpacityTransferFunction.linear(lo, hi), size)


def c1_scene(size=128, normals=None, bounds=0.5):
    """Config C1: double_sphere + normals, bounds +-0.5, directional light, viridis + linear(0, 0.3)."""
    data = create_sample_volume(size, "double_sphere")
    b = np.float32(bounds)
    vol = Volume(data=data, normals=normals,
                 min_bounds=np.array([-b, -b, -b], np.float32), max_bounds=np.array([b, b, b], np.float32))
    return vol, Light.directional([1, -1, 0]), viridis_lut()


def turntable_camera(k, n=360):
    """Config C3 view k: azimuth 2*pi*k/n, elevation pi/6, distance 3."""
    return Camera.from_spherical(target=np.array([0.0, 0.0, 0.0], dtype=np.float32),
                                 azimuth=2 * np.pi * k / n, elevation=np.pi / 6, roll=0.0, distance=3.0)


def image_metrics(got, want):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    mse = float(np.mean((got.astype(np.float64) - want.astype(np.float64)) ** 2))
    return {
        "max_abs": int(d.max()),
        "frac_within_2": float((d <= 2).all(axis=-1).mean()),
        "frac_within_1": float((d <= 1).all(axis=-1).mean()),
        "frac_identical": float((d == 0).all(axis=-1).mean()),
        "psnr_db": float("inf") if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse)),
    }


def assert_parity(got, want, min_frac=0.999, min_psnr=45.0):
    """The tolerance BASELINE.json states: |delta| <= 2/255 on >= 99.9 % of pixels, PSNR >= 45 dB."""
    m = image_metrics(got, want)
    assert m["frac_within_2"] >= min_frac and m["psnr_db"] >= min_psnr, m
    return m
