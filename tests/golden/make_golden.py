"""Generate tests/golden/* by importing the REFERENCE's own Python (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

The reference package eagerly imports moderngl / vtk / matplotlib (pyvr/__init__.py:10-19),
none of which exist in this image, so those three are stubbed with MagicMock modules; every
value written here comes from the reference's numpy-only code (datasets, camera, lighting,
config, transferfunctions minus from_colormap).  The GPU box has no /root/reference: tests read
only the files this script wrote.

Outputs:
  camera.json       get_camera_vectors / view / projection for presets and odd parameter sets
  lights.json       Light presets -> position/target/intensities
  presets.json      RenderConfig presets
  luts.npz          ColorTransferFunction / OpacityTransferFunction .to_lut outputs
  volumes.npz       create_sample_volume(16, shape) for the five deterministic analytic shapes
  normals.npz       compute_normal_volume inputs/outputs (cubic, ragged, tiny, constant)
  meta.json         sha256 of the shader sources the oracle restates + of larger arrays
"""

import hashlib
import importlib.abc
import importlib.machinery
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    TOP = ("moderngl", "vtk", "matplotlib")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.TOP:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _StubFinder())
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from pyvr.camera import Camera  # noqa: E402
from pyvr.config import RenderConfig  # noqa: E402
from pyvr.datasets import compute_normal_volume, create_sample_volume  # noqa: E402
from pyvr.lighting import Light  # noqa: E402
from pyvr.transferfunctions import ColorTransferFunction, OpacityTransferFunction  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def file_sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


# ---------------------------------------------------------------- cameras
def cam_record(name, cam, aspects=(1.0, 16 / 9)):
    pos, up = cam.get_camera_vectors()
    return {
        "name": name,
        "params": cam.to_dict(),
        "position": np.asarray(pos, dtype=np.float64).tolist(),
        "up": np.asarray(up, dtype=np.float64).tolist(),
        "view": cam.get_view_matrix().astype(np.float64).tolist(),
        "proj": {repr(float(a)): cam.get_projection_matrix(a).astype(np.float64).tolist() for a in aspects},
    }


f32 = lambda *v: np.array(v, dtype=np.float32)  # noqa: E731
cams = [
    ("default", Camera()),
    ("front", Camera.front_view(distance=3.0)),
    ("side", Camera.side_view(distance=3.0)),
    ("top", Camera.top_view(distance=3.0)),
    ("iso", Camera.isometric_view(distance=3.0)),
    ("iso_far", Camera.isometric_view(distance=7.5)),
    ("benchmark", Camera.from_spherical(target=np.array([0.0, 0.0, 0.0]), distance=3.0,
                                        azimuth=np.pi / 4, elevation=np.pi / 6, roll=0.0)),
    ("rolled", Camera(target=f32(0.1, -0.2, 0.3), azimuth=1.1, elevation=-0.4, roll=0.7, distance=2.2)),
    ("yup", Camera(azimuth=0.3, elevation=0.2, roll=-0.1, distance=4.0,
                   init_pos=f32(0, 0, 1), init_up=f32(0, 1, 0), fov=np.radians(60))),
    ("odd_vectors", Camera(target=f32(0.5, 0.5, 0.0), azimuth=-2.0, elevation=1.2, roll=3.0,
                           distance=1.5, init_pos=f32(2, 1, 0.5), init_up=f32(0, 0, 2), fov=0.5)),
]
for k in range(0, 360, 45):
    cams.append((f"turntable_{k:03d}",
                 Camera.from_spherical(target=np.array([0.0, 0.0, 0.0]), azimuth=2 * np.pi * k / 360,
                                       elevation=np.pi / 6, roll=0.0, distance=3.0)))
with open(os.path.join(OUT, "camera.json"), "w") as f:
    json.dump([cam_record(n, c) for n, c in cams], f, indent=1)

# ---------------------------------------------------------------- lights / presets
iso = Camera.isometric_view(distance=3.0)
linked = Light.camera_linked()
linked.update_from_camera(iso)
lights = {"default": Light.default(), "directional_1m10": Light.directional([1, -1, 0]),
          "directional_custom": Light.directional(np.array([0.2, 0.5, -1.0]), ambient=0.3, diffuse=0.6, distance=4.0),
          "point": Light.point_light([5, 5, 5], ambient=0.1, diffuse=0.7),
          "ambient_only": Light.ambient_only(0.3), "camera_linked_iso": linked}
with open(os.path.join(OUT, "lights.json"), "w") as f:
    json.dump({k: {"position": np.asarray(v.position, np.float64).tolist(),
                   "target": np.asarray(v.target, np.float64).tolist(),
                   "ambient": v.ambient_intensity, "diffuse": v.diffuse_intensity,
                   "direction": np.asarray(v.get_direction(), np.float64).tolist()}
               for k, v in lights.items()}, f, indent=1)

with open(os.path.join(OUT, "presets.json"), "w") as f:
    json.dump({name: {"step_size": c.step_size, "max_steps": c.max_steps,
                      "early_ray_termination": c.early_ray_termination,
                      "opacity_threshold": c.opacity_threshold,
                      "reference_step_size": c.reference_step_size,
                      "samples_per_ray": c.estimate_samples_per_ray(),
                      "relative_time": c.estimate_render_time_relative(), "repr": repr(c)}
               for name, c in (("preview", RenderConfig.preview()), ("fast", RenderConfig.fast()),
                               ("balanced", RenderConfig.balanced()),
                               ("high_quality", RenderConfig.high_quality()),
                               ("ultra_quality", RenderConfig.ultra_quality()),
                               ("default", RenderConfig()))}, f, indent=1)

# ---------------------------------------------------------------- LUTs
ctf_pts = [(0.0, (0.0, 0.0, 0.2)), (0.25, (0.1, 0.9, 0.3)), (0.6, (1.0, 0.5, 0.0)), (1.0, (1.0, 1.0, 1.0))]
luts = {
    "otf_linear_0_1": OpacityTransferFunction.linear(0.0, 1.0).to_lut(),
    "otf_linear_0_0p3": OpacityTransferFunction.linear(0.0, 0.3).to_lut(),
    "otf_linear_0_0p1_64": OpacityTransferFunction.linear(0.0, 0.1).to_lut(64),
    "otf_one_step": OpacityTransferFunction.one_step(0.5, 0.0, 0.8).to_lut(),
    "otf_peaks": OpacityTransferFunction.peaks([0.3, 0.7], opacity=0.9, eps=0.05, base=0.1).to_lut(),
    "otf_custom": OpacityTransferFunction([(0.0, 0.0), (0.3, 0.1), (0.8, 0.9), (1.0, 0.5)]).to_lut(100),
    "ctf_gray": ColorTransferFunction.grayscale().to_lut(),
    "ctf_custom": ColorTransferFunction(ctf_pts).to_lut(),
    "ctf_custom_17": ColorTransferFunction(ctf_pts).to_lut(17),
    "ctf_single": ColorTransferFunction.single_color((0.2, 0.4, 0.6)).to_lut(32),
}
np.savez_compressed(os.path.join(OUT, "luts.npz"), **luts)

# ---------------------------------------------------------------- sample volumes
shapes = ["sphere", "torus", "double_sphere", "cube", "helix"]
np.savez_compressed(os.path.join(OUT, "volumes.npz"), **{s: create_sample_volume(16, s) for s in shapes})
vol_sha = {f"{s}_{n}": sha(create_sample_volume(n, s)) for s in shapes for n in (33, 64)}
vol_sha["random_blob_32"] = sha(create_sample_volume(32, "random_blob"))

# ---------------------------------------------------------------- normals
rng = np.random.default_rng(20261017)
cases = {
    "double_sphere_24": create_sample_volume(24, "double_sphere"),
    "cube_20": create_sample_volume(20, "cube"),               # large exactly-zero regions
    "ragged_9x12x17": rng.random((9, 12, 17), dtype=np.float32),
    "ragged_31x2x5": (rng.standard_normal((31, 2, 5)) * 100).astype(np.float32),
    "tiny_2x2x2": rng.random((2, 2, 2), dtype=np.float32),
    "constant_6": np.full((6, 6, 6), 0.25, dtype=np.float32),
    "denormalish_8": (rng.random((8, 8, 8), dtype=np.float32) * 1e-9).astype(np.float32),
}
npz = {}
for k, v in cases.items():
    npz[k + "__in"] = v
    npz[k + "__out"] = compute_normal_volume(v)
    assert npz[k + "__out"].dtype == np.float32
np.savez_compressed(os.path.join(OUT, "normals.npz"), **npz)
nrm_sha = {"double_sphere_128": sha(compute_normal_volume(create_sample_volume(128, "double_sphere")))}

# ---------------------------------------------------------------- meta
meta = {
    "reference_version": "0.4.1",
    "shader_sha256": {
        "pyvr/shaders/volume.frag.glsl": file_sha(os.path.join(REF, "pyvr/shaders/volume.frag.glsl")),
        "pyvr/shaders/volume.vert.glsl": file_sha(os.path.join(REF, "pyvr/shaders/volume.vert.glsl")),
    },
    "example_data_sha256": {n: file_sha(os.path.join(REF, "example_data", n)) for n in ("fuel.vti", "hydrogen.vti")},
    "sample_volume_sha256": vol_sha,
    "normal_volume_sha256": nrm_sha,
    "numpy": np.__version__,
}
with open(os.path.join(OUT, "meta.json"), "w") as f:
    json.dump(meta, f, indent=1)
print("golden written to", OUT)
