#!/usr/bin/env python
"""Headline benchmark: Gsamples/s of the volume ray march on BASELINE.json's config C3.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (SURVEY.md section 8 d, C3 = BASELINE.json configs[2]): synthetic 512^3 float32
``double_sphere`` volume + normal volume, bounds +-1, 1920x1080, ``RenderConfig.high_quality()``
(step 0.005, 1000 steps, reference step 0.01), ``Light.directional([1,-1,0])``, viridis colour TF +
``linear(0, 0.1)`` opacity TF, the 360-view turntable (azimuth 2*pi*k/360, elevation pi/6, distance 3).

A *step* is one batch of ``--views-per-step`` turntable views per GPU (default 180 = half a revolution, so
that the driver's 20 timed steps last about two seconds).  Views are dealt ``k mod N`` over the N ranks (no
data-path collective; every rank holds the whole volume), so per-GPU work is fixed as N grows:
``"scaling": "weak"``.  A *sample* is one in-box loop body of the reference shader (volume.frag.glsl:92-116),
counted by the kernel; ``value`` = samples of all ranks / max-over-ranks device time, frames
device-resident.  ``e2e`` = the same through ``VolumeRenderer.render_batch`` with the views coming from
pinned host memory and the RGBA8 frames read back to pinned host memory inside the timed region.

``--impl reference`` times the reference's own shader (pyvr/shaders/volume.frag.glsl, verbatim) on Mesa
llvmpipe on the host cores -- the baseline BASELINE.json names -- through ``oracle/gl`` (a ctypes replay of
pyvr/moderngl_renderer/manager.py; moderngl itself is not in the image): ``"kind": "reference"``.  One step =
one whole 1920x1080 view of the same turntable (a bounded sample of the GPU arm's step).  If the Mesa
library is missing it falls back to the C/OpenMP restatement (``oracle/``, ``"kind": "port"``).

At N > 1 the line also carries ``secondary`` (time-boxed C4 image-tile and C5 sort-last lines, see
bench_partitioned.py) and ``parity_check`` (tile-sharded frame == single-GPU frame bit for bit, bricked
composite vs single GPU within tolerance, relay bit-identical), so that the multi-process paths are
measured and checked under the driver.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_SAMPLE_F32 = 128   # 8 texels x {s,nx,ny,nz} binary32 (SURVEY.md section 8 d)
BYTES_PER_SAMPLE_F16 = 64


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 = headline turntable (default); c4 = image tiles over a replicated device-generated "
                         "volume; c5 = sort-last bricks + binary swap (secondary lines, see DESIGN.md)")
    ap.add_argument("--size", type=int, default=None, help="volume edge (default 512 / 2048 / 2048*cbrt(N))")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="c5: binary-swap exchange path")
    ap.add_argument("--views-per-step", type=int, default=180)
    ap.add_argument("--texels", default="f32", choices=["f32", "f16"])
    ap.add_argument("--no-ess", action="store_true", help="disable empty-space skipping")
    ap.add_argument("--hwtex", action="store_true", help="sample through the texture unit (hardware trilinear)")
    ap.add_argument("--layout", default=None, choices=["linear", "swizzle"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-alternatives", action="store_true", help="skip the f16 / hwtex / dense side measurements")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C4 (N = 1) / C4 + C5 + parity check (N > 1) section")
    ap.add_argument("--secondary-seconds", type=float, default=240.0, help="N > 1: time box of the secondary section")
    ap.add_argument("--reference-backend", default="auto", choices=["auto", "gl", "oracle"],
                    help="CPU arm: gl = the reference's GLSL on Mesa llvmpipe; oracle = C/OpenMP restatement")
    return ap.parse_args()


def scene(size):
    """C3 inputs on the host (volume without normals; the normal volume is computed on the GPU)."""
    from pyvr_b200 import (ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                           build_rgba_lut, create_sample_volume)

    data = create_sample_volume(size, "double_sphere")
    light = Light.directional([1, -1, 0])
    config = RenderConfig.high_quality()
    lut = build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.1))
    return data, light, config, lut


def turntable_camera(k, n=360):
    from pyvr_b200 import Camera

    return Camera.from_spherical(target=np.array([0.0, 0.0, 0.0], dtype=np.float32),
                                 azimuth=2 * np.pi * (k % n) / n, elevation=np.pi / 6, roll=0.0, distance=3.0)


def step_view_indices(step, rank, world, per_step):
    """View k of the turntable goes to rank k mod world."""
    base = step * per_step * world
    return [base + j * world + rank for j in range(per_step)]


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region: NVML in-process every 5 ms
    (nvidia_ml_py), falling back to polling the nvidia-smi binary every 200 ms."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self._stop, self._thread = index, [], threading.Event(), None
        self.source, self._nvml, self._handle, self.sm_max = "nvidia-smi", None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml, self.source = pynvml, "nvml"
        except Exception:
            self._nvml = None

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    mhz = float(self._nvml.nvmlDeviceGetClockInfo(self._handle, self._nvml.NVML_CLOCK_SM))
                    try:
                        mask = int(self._nvml.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
                    except Exception:
                        mask = int(self._nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
                    self.rows.append([mhz, self.sm_max] + [bool(mask & self.NVML_BITS[n]) for n in self.NAMES])
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    if out.returncode == 0 and out.stdout.strip():
                        c = [x.strip() for x in out.stdout.strip().split(",")]
                        self.rows.append([float(c[0]), float(c[1])] + [x.lower().startswith("active") for x in c[2:6]])
            except Exception:
                pass
            self._stop.wait(0.005 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < 0.5:     # first sample in hand before the timed region starts
            time.sleep(0.001)
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2 + i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core.  Must run
    before the OpenMP runtime of liboracle.so initialises and before llvmpipe creates its rasteriser threads."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["LP_NUM_THREADS"] = str(min(n, 16))     # Mesa 18 caps llvmpipe at 16 rasteriser threads
    return n


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's shader on Mesa llvmpipe (oracle/gl), else the C/OpenMP restatement (oracle/)
# --------------------------------------------------------------------------------------------------
LLVMPIPE_CHECK_SIZE = 384   # largest multiple of 128 whose RGB32F normal texture fits Mesa 18 llvmpipe's 1 GiB cap (406^3)


class CpuReference:
    """C3 scene loaded once into the CPU renderer; ``render_view(k)`` = one whole turntable view.

    Preferred: the reference's own GLSL on Mesa llvmpipe (``oracle/gl``, ``kind = "reference"``).  The Mesa 18.1.9
    build in this image caps a texture at 1 GiB and stores RGB32F as RGBA32F, so the 512^3 normal texture of C3
    (2 GiB) cannot be created: at that size the arm falls back to the C/OpenMP restatement (``kind = "port"``) and
    ``llvmpipe_check()`` times both CPU renderers on the same scene at 384^3 to show how they compare."""

    def __init__(self, args, data, normals, light, config, lut, backend="auto"):
        import oracle
        from pyvr_b200 import Volume

        self.args, self.light, self.config, self.lut = args, light, config, lut
        self.vol = Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32),
                          max_bounds=np.array([1, 1, 1], np.float32))
        self.oracle, self.gl, self.kind = oracle, None, "port"
        self.cores = oracle.num_threads()
        self.what = "CPU restatement of the reference shader (oracle/pyvr_oracle.c, C + OpenMP)"
        self.gl_error = None
        if backend in ("auto", "gl"):
            try:
                self.gl = self._gl_renderer(self.vol)
                self.kind = "reference"
                self.cores = int(os.environ.get("LP_NUM_THREADS", min(host_threads(), 16)))
                self.what = ("the reference's own GLSL (pyvr/shaders/volume.frag.glsl, verbatim) on " + self.gl.info["renderer"]
                             + ", " + self.gl.info["version"] + ", driven as pyvr/moderngl_renderer/manager.py drives it "
                             "(oracle/gl: ctypes OpenGL, moderngl is not in the image); clear + draw + glReadPixels per view")
            except Exception as e:   # GLUnavailable (no Mesa, texture too large), compile errors, ...
                if backend == "gl":
                    raise
                self.gl_error = f"{type(e).__name__}: {e}"
                self.what += "; the reference's GLSL on Mesa llvmpipe could not take this scene: " + self.gl_error

    def _gl_renderer(self, vol):
        import oracle.gl as ogl

        r = ogl.GLReference(self.args.width, self.args.height)
        r.load_shaders()
        r.set_config(self.config.step_size, self.config.max_steps, self.config.reference_step_size)
        r.set_light(self.light.ambient_intensity, self.light.diffuse_intensity, self.light.position, self.light.target)
        r.load_volume(vol.data, vol.normals, vol.min_bounds, vol.max_bounds)
        r.set_lut(self.lut)
        return r

    def _gl_view(self, gl, k):
        cam = turntable_camera(k)
        pos, _ = cam.get_camera_vectors()
        gl.set_camera(cam.get_view_matrix(), cam.get_projection_matrix(self.args.width / self.args.height), pos)
        t0 = time.perf_counter()
        gl.render()
        return time.perf_counter() - t0

    def samples_of(self, k):
        """Reference samples of view k (the unit of the metric): counted by the restatement, which executes the
        same loop bodies as the shader (tests/test_gl_reference.py)."""
        _, _, st = self.oracle.render(self.vol, turntable_camera(k), self.light, self.config, self.lut,
                                      self.args.width, self.args.height)
        return st["samples"]

    def render_view(self, k):
        """Returns (seconds, samples or None).  With GL the sample count comes from ``samples_of`` (untimed)."""
        if self.gl is not None:
            return self._gl_view(self.gl, k), None
        cam = turntable_camera(k)
        t0 = time.perf_counter()
        _, _, st = self.oracle.render(self.vol, cam, self.light, self.config, self.lut, self.args.width, self.args.height)
        return time.perf_counter() - t0, st["samples"]

    def llvmpipe_check(self, k=0):
        """The reference's GLSL on llvmpipe next to the C restatement on the SAME scene at 384^3 (the largest
        volume this llvmpipe can hold), one whole view each after a warm-up view."""
        from pyvr_b200 import Volume, create_sample_volume

        try:
            data = create_sample_volume(LLVMPIPE_CHECK_SIZE, "double_sphere")
            vol = Volume(data=data, normals=self.oracle.normals(data), min_bounds=self.vol.min_bounds, max_bounds=self.vol.max_bounds)
            gl = self._gl_renderer(vol)
            self._gl_view(gl, k)
            t_gl = self._gl_view(gl, k)
            info = gl.info
            gl.close()
            cam = turntable_camera(k)
            t0 = time.perf_counter()
            _, _, st = self.oracle.render(vol, cam, self.light, self.config, self.lut, self.args.width, self.args.height)
            t_port = time.perf_counter() - t0
            return {"scene": f"C3 with a {LLVMPIPE_CHECK_SIZE}^3 volume, {self.args.width}x{self.args.height}, turntable view {k}",
                    "samples": st["samples"], "llvmpipe_Gsamples_per_s": st["samples"] / t_gl / 1e9, "llvmpipe_seconds": t_gl,
                    "llvmpipe_threads": int(os.environ.get("LP_NUM_THREADS", min(host_threads(), 16))),
                    "port_Gsamples_per_s": st["samples"] / t_port / 1e9, "port_seconds": t_port,
                    "port_threads": self.oracle.num_threads(), "port_over_llvmpipe": t_gl / t_port,
                    "gl": info["renderer"] + ", " + info["version"]}
        except Exception as e:
            return {"error": f"{type(e).__name__}: {e}"}


def cpu_baseline(ref, view_indices, target_seconds):
    """Whole views of the step, one after the other, until `target_seconds` of CPU time have been spent."""
    ref.render_view(view_indices[0])            # warm-up: shader JIT, page-in (the reference's protocol has one too)
    samples, views, dt = 0, 0, 0.0
    for k in view_indices:
        sec, n = ref.render_view(k)
        samples += n if n is not None else ref.samples_of(k)
        dt += sec
        views += 1
        if dt >= target_seconds:
            break
    out = {
        "value": samples / dt / 1e9, "unit": "Gsamples/s", "cores": ref.cores, "kind": ref.kind,
        "sample": f"{views} whole views of the step (turntable views {view_indices[0]}..{view_indices[views - 1]}, "
                  f"{samples} samples) in {dt:.2f} s after one warm-up view; {ref.what}",
        "seconds": dt, "frames_per_s": views / dt,
    }
    if ref.kind == "port":
        out["llvmpipe_check"] = ref.llvmpipe_check(view_indices[0])
    return out


def run_reference(args, rank):
    """--impl reference: the CPU path on rank 0 only."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    import oracle

    data, light, config, lut = scene(args.size)
    normals = oracle.normals(data)   # the reference's compute_normal_volume, restated (bit-exact vs numpy)
    ref = CpuReference(args, data, normals, light, config, lut, args.reference_backend)
    per_step = args.views_per_step
    # one step = the first whole view of the GPU arm's step (a bounded sample: 1 of `per_step` views)
    ks = [step_view_indices(s, 0, 1, per_step)[0] for s in range(args.warmup + args.steps)]
    for k in ks[:args.warmup]:
        ref.render_view(k)
    dt, samples = 0.0, 0
    t_wall = time.perf_counter()
    for k in ks[args.warmup:]:
        sec, n = ref.render_view(k)
        dt += sec
        samples += n if n is not None else 0
    wall = time.perf_counter() - t_wall
    if ref.gl is not None:                      # counted outside the timed region
        samples = sum(ref.samples_of(k) for k in ks[args.warmup:])
    value = samples / wall / 1e9
    sample = (f"1 whole view of each step's {per_step} (turntable views {ks[args.warmup]}, {ks[args.warmup] + per_step}, ...); "
              + ref.what)
    line = {
        "impl": "reference", "metric": "ray-march throughput", "value": value, "unit": "Gsamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "frames_per_s": args.steps / wall,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": ref.cores, "kind": ref.kind, "sample": sample,
                         "host_cores": cores, **({"llvmpipe_check": ref.llvmpipe_check(ks[args.warmup])} if ref.kind == "port" else {})},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {
        "workload": f"C3: synthetic {args.size}^3 f32 double_sphere + normals, bounds +-1, {args.width}x{args.height}, "
                    "high_quality (step 0.005, 1000 steps), 360-view turntable (el 30 deg, d 3), "
                    "directional light, viridis + linear(0,0.1)",
        "views_per_step_per_gpu": args.views_per_step,
        "sharding": "view k -> rank k mod N, volume replicated, no collective",
        "texels": "f32x4 (16 B/voxel)" if args.texels == "f32" else "f16x4 (8 B/voxel)",
        "l2_policy": f"inputs larger than L2 (packed volume {args.size ** 3 * (16 if args.texels == 'f32' else 8) / 2 ** 30:.2f} GiB vs 126 MB L2), no flush",
        "empty_space_skipping": not args.no_ess,
        "sampling": "texture unit, hardware trilinear (8-bit weights)" if getattr(args, "hwtex", False) else "binary32 software trilinear",
    }


# --------------------------------------------------------------------------------------------------
# K2: compute_normal_volume
# --------------------------------------------------------------------------------------------------
def numpy_normals(volume):
    """The reference's compute_normal_volume (pyvr/datasets/synthetic.py:118-122) as numpy evaluates it."""
    grad = np.stack(np.gradient(volume), axis=-1)
    return (grad / (np.linalg.norm(grad, axis=-1, keepdims=True) + 1e-8)).astype(np.float32)


def time_normals(torch, _cabi, data, device, peak):
    """K2: device-resident kernel time (best of 5 after a warm-up, CUDA events inside the C ABI call) and the
    host-array-in / host-array-out call a user of the reference makes, next to the reference's numpy function."""
    n0, n1, n2 = data.shape
    d_in = torch.from_numpy(data).cuda(device)
    d_out = torch.empty((n0, n1, n2, 3), dtype=torch.float32, device=d_in.device)
    ms, best = ctypes.c_float(0.0), float("inf")
    for _ in range(6):
        _cabi.check(_cabi.lib().pyvr_cuda_compute_normals(device, ctypes.c_void_p(d_in.data_ptr()),
                                                          ctypes.c_void_p(d_out.data_ptr()), n0, n1, n2, 1, ctypes.byref(ms)))
        best = min(best, ms.value)
    # opt-in PYVR_NORMALS_RELAXED (g * (1/norm) instead of correctly rounded quotients): time and worst error
    exact = d_out.clone()
    best_relaxed = float("inf")
    for _ in range(6):
        _cabi.check(_cabi.lib().pyvr_cuda_compute_normals(device, ctypes.c_void_p(d_in.data_ptr()),
                                                          ctypes.c_void_p(d_out.data_ptr()), n0, n1, n2, 1 | 2, ctypes.byref(ms)))
        best_relaxed = min(best_relaxed, ms.value)
    relaxed_err = float(((d_out - exact).abs() / exact.abs().clamp_min(1.0)).max().item())
    relaxed_ulp = int((d_out.view(torch.int32) - exact.view(torch.int32)).abs().max().item())
    del d_in, d_out, exact
    torch.cuda.empty_cache()
    e2e = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        normals = _cabi.compute_normals_host(data, device=device)
        e2e = min(e2e, time.perf_counter() - t0)
    t0 = time.perf_counter()
    want = numpy_normals(data)
    t_numpy = time.perf_counter() - t0
    ok = bool(np.array_equal(want.view(np.uint32), normals.view(np.uint32)))
    gbs = data.size * 16 / (best * 1e-3) / 1e9
    mix_gbs = _cabi.measure_cache_bandwidth(3, device)      # DRAM rate of a 1:3 read:write stream, measured now
    return normals, {
        "kernel_ms": best, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak, "algorithmic_bytes_per_voxel": 16,
        "mix_1r3w_peak_GB/s": mix_gbs, "frac_of_mix_peak": gbs / mix_gbs if mix_gbs else None,
        "mix_peak_source": "measured live: pyvr_cuda_measure_cache_bandwidth(level 3) -- streaming kernel with this stencil's "
                           "traffic mix (4 B read + 12 B written per element, no arithmetic), csrc/bandwidth.cu",
        "e2e": {"seconds": e2e, "Mvoxels/s": data.size / e2e / 1e6, "api": "compute_normal_volume(host array) -> host array",
                "h2d_bytes": data.nbytes, "d2h_bytes": data.nbytes * 3},
        "cpu_baseline": {"seconds": t_numpy, "Mvoxels/s": data.size / t_numpy / 1e6, "kind": "reference",
                         "what": "numpy restatement of pyvr/datasets/synthetic.py:118-122 (np.gradient, stack, norm), one process"},
        "e2e_speedup_vs_numpy": t_numpy / e2e, "bit_identical_to_numpy": ok,
        "relaxed_opt_in": {"kernel_ms": best_relaxed, "GB/s": data.size * 16 / (best_relaxed * 1e-3) / 1e9,
                           "frac_of_hbm_peak": data.size * 16 / (best_relaxed * 1e-3) / 1e9 / peak,
                           "max_abs_err_over_max(1,|ref|)": relaxed_err, "max_ulp_distance": relaxed_ulp,
                           "what": "pyvr_cuda_compute_normals(..., PYVR_NORMALS_RELAXED): quotients as g * (1/norm); "
                                   "tolerance of the path 1e-5; NOT the default, which stays bit-identical to numpy"},
    }


def active_texel_bytes(data, lut, entry_bytes, cell=4):
    """Bytes of packed texels in ACTIVE macrocells (`cell`^3 voxels whose scalar range can reach a non-zero LUT alpha;
    4 = the library's PYVR_CELL_SHIFT 2): what the ESS march can touch per frame (SURVEY.md section 8 d, 'unique brick
    bytes touched')."""
    n = data.shape[0]
    c = n // cell
    if n % cell or data.shape != (n, n, n):
        return None
    hi = data.reshape(c, cell, c, cell, c, cell).max(axis=(1, 3, 5))
    for axis in range(3):                                  # +1 apron the upper taps reach
        nb = np.concatenate([np.take(hi, range(1, c), axis=axis), np.take(hi, [c - 1], axis=axis)], axis=axis)
        hi = np.maximum(hi, nb)
    size = lut.shape[0]
    nz = np.concatenate([[0], np.cumsum(lut[:, 3] != 0)])
    jh = np.clip(np.floor(hi.astype(np.float64) * size - 0.5) + 1, 0, size - 1).astype(np.int64)
    active = nz[jh + 1] > 0                                # lower end of every cell of this volume is ~0
    return int(active.sum()) * cell ** 3 * entry_bytes, float(active.mean())


def traffic_per_view(args):
    """DRAM bytes one view of the march moves, from the committed ncu capture (profiles/r02_traffic.json)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            table = json.load(open(os.path.join(ROOT, "profiles", name)))
            key = f"{args.texels}_{'hwtex_' if args.hwtex else ''}{'dense' if args.no_ess else 'ess'}"
            return float(table["c3_dram_bytes_per_view"][key]), name
        except Exception:
            continue
    return None, None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload != "c3":
        import bench_partitioned

        bench_partitioned.run(args, rank, world, local_rank)
        return
    args.size = args.size or 512
    args.width, args.height = args.width or 1920, args.height or 1080
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyvr_b200 has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.layout:
        os.environ["PYVR_CUDA_LAYOUT"] = args.layout

    from pyvr_b200 import Volume
    from pyvr_b200.cuda_renderer import VolumeRenderer, _cabi

    peak, peak_src = measured_peak_gbs()
    t_setup = time.perf_counter()
    renderer_kw = dict(device=local_rank, texel_format=args.texels, empty_space_skipping=not args.no_ess,
                       hardware_filtering=args.hwtex)
    normals_info = None
    if world == 1:
        # host pipeline, as a user of the reference would: create_sample_volume -> compute_normal_volume (K2 on
        # the GPU) -> Volume -> load_volume.  The host arrays also feed the CPU baseline.
        data, light, config, lut = scene(args.size)
        normals, normals_info = time_normals(torch, _cabi, data, local_rank, peak)
        vol = Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32),
                     max_bounds=np.array([1, 1, 1], np.float32))
        renderer = VolumeRenderer(args.width, args.height, config=config, light=light, **renderer_kw)
        renderer.load_volume(vol)
    else:
        # N > 1: every rank builds the SAME texels on its own device (synth.cu; identical to the host pipeline,
        # tests/test_synth_gpu.py) instead of N processes each holding ~6 GB of host temporaries
        from pyvr_b200 import (ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig, build_rgba_lut)

        data = normals = vol = None
        light, config = Light.directional([1, -1, 0]), RenderConfig.high_quality()
        lut = build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.1))
        renderer = VolumeRenderer(args.width, args.height, config=config, light=light, **renderer_kw)
        normals_info = {"device_generation_ms": renderer.generate_volume(args.size, "double_sphere", (-1, -1, -1), (1, 1, 1))}
    renderer.set_lut(lut)
    texel_layout = renderer.texel_layout
    stream = torch.cuda.Stream()   # non-default: the library treats stream 0 as "use the context's own stream"
    renderer.set_stream(stream.cuda_stream)
    setup_s = time.perf_counter() - t_setup

    # roofline denominators measured on this GPU, now (csrc/bandwidth.cu)
    l1_gbs = _cabi.measure_cache_bandwidth(1, local_rank)
    l2_gbs = _cabi.measure_cache_bandwidth(2, local_rank)

    per_step = args.views_per_step
    frame_bytes = args.width * args.height * 4
    d_frames = torch.empty(per_step * frame_bytes, dtype=torch.uint8, device="cuda")
    pinned = _cabi.PinnedBuffer(per_step * frame_bytes)
    total_steps = args.warmup + args.steps
    views = [renderer.make_views(turntable_camera(k) for k in step_view_indices(s, rank, world, per_step))
             for s in range(total_steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- device-resident pass: `value` + roofline -----------------------------------
    def resident_step(s, r=renderer):
        r.render_batch(views=views[s], device_ptr=d_frames.data_ptr())
        return r.stats

    for s in range(args.warmup):
        resident_step(s)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = fetched = launches = 0
    kernel_ms = 0.0
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for s in range(args.warmup, total_steps):
            st = resident_step(s)
            samples += st["samples"]
            fetched += st["samples_fetched"]
            launches += st["kernel_launches"]
            kernel_ms += st["kernel_ms"]
        e1.record(stream)
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    all_samples = sum_over_ranks(float(samples))
    value = all_samples / (ms * 1e-3) / 1e9
    frames = args.steps * per_step * world

    # ---------------- end-to-end pass: views from pinned host memory, frames back to pinned host memory
    for s in range(args.warmup):
        renderer.render_batch(views=views[s], out=pinned.array)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    e2e_samples = 0
    for s in range(args.warmup, total_steps):
        renderer.render_batch(views=views[s], out=pinned.array)
        e2e_samples += renderer.stats["samples"]
    e1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), e2e_wall_ms))
    e2e_value = sum_over_ranks(float(e2e_samples)) / (e2e_ms * 1e-3) / 1e9
    checksum = int(pinned.array[::4099].astype(np.uint64).sum())

    # ---------------- side measurements (N = 1 only), a few steps each, device-resident:
    # the dense march (no empty-space skipping) for the roofline of the fetch path itself, and the storage /
    # sampler alternatives the north star allows when within tolerance.  Not the headline: the headline is
    # binary32 texels + binary32 software trilinear + exact skipping.
    def side_run(kw, steps):
        r = VolumeRenderer(args.width, args.height, config=config, light=light, device=local_rank, **kw)
        r.load_volume(vol)
        r.set_lut(lut)
        r.set_stream(stream.cuda_stream)
        resident_step(0, r)
        torch.cuda.synchronize()
        e0.record(stream)
        n_s = n_f = n_l = 0
        k_ms = 0.0
        for s_ in range(steps):
            st_ = resident_step(args.warmup + s_, r)
            n_s, n_f, n_l, k_ms = n_s + st_["samples"], n_f + st_["samples_fetched"], n_l + st_["kernel_launches"], k_ms + st_["kernel_ms"]
        e1.record(stream)
        torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1)
        r.close()
        return {"value": n_s / (t_ms * 1e-3) / 1e9, "unit": "Gsamples/s", "frames_per_s": steps * per_step / (t_ms * 1e-3),
                "fetched_Gsamples/s": n_f / (k_ms * 1e-3) / 1e9, "kernel_ms_per_view": k_ms / (steps * per_step)}

    alternatives, dense = {}, None
    if world == 1 and not args.no_alternatives and not args.hwtex and args.texels == "f32" and not args.no_ess:
        side_steps = max(1, min(args.steps, 3))
        dense = side_run(dict(texel_format="f32", empty_space_skipping=False), 1)
        for name, kw in (("f16x4 texels, software trilinear", dict(texel_format="f16")),
                         ("f16x4 texels, texture-unit trilinear (hwtex)", dict(texel_format="f16", hardware_filtering=True))):
            alternatives[name] = side_run(dict(empty_space_skipping=True, **kw), side_steps)

    line = None
    if rank == 0:
        bytes_per_sample = BYTES_PER_SAMPLE_F32 if args.texels == "f32" else BYTES_PER_SAMPLE_F16
        launch_ms = kernel_ms / max(launches, 1)
        fetched_per_launch = fetched / max(launches, 1)
        achieved = fetched_per_launch * bytes_per_sample / (launch_ms * 1e-3) / 1e9
        views_per_launch = per_step / max(launches / max(args.steps, 1), 1)
        per_view, traffic_file = traffic_per_view(args)
        sm_mhz = clocks.summary().get("sm_mhz") or 1965.0
        l1_nominal = 148 * 128 * sm_mhz * 1e6 / 1e9      # one 128-byte load-return wavefront per SM per clock
        frames_per_s_kernel = views_per_launch / (launch_ms * 1e-3)
        hbm = None
        if data is not None:
            texel_b = 16 if args.texels == "f32" else 8
            entry_b = texel_b * (2 if "z-pair" in texel_layout else 1)
            got = active_texel_bytes(data, lut, entry_b)
            if got is not None:
                unique = got[0] if not args.no_ess else data.size * entry_b
                hbm_achieved = (unique + frame_bytes) * frames_per_s_kernel / 1e9
                hbm = {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                       "unique_texel_bytes_per_frame": unique, "active_cell_fraction": got[1], "peak_source": peak_src,
                       "note": "SURVEY.md section 8(d): (unique brick bytes touched per frame + W*H*4) x frames/s of the kernel / "
                               f"HBM copy rate; unique = packed entries ({entry_b} B per voxel, layout: {texel_layout}) of the active 4^3 macrocells"}
        line = {
            "metric": "ray-march throughput", "value": value, "unit": "Gsamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "frames_per_s": frames / (ms * 1e-3),
            "samples_per_frame": all_samples / frames,
            "timed_region_s": ms * 1e-3,
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "frames_per_s": frames / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": world * per_step * ctypes.sizeof(_cabi.View),
                    "d2h_bytes_per_step": world * per_step * frame_bytes,
                    "api": "VolumeRenderer.render_batch(views, out=pinned host buffer)", "frame_checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {
                # the binding unit of this gather kernel is the SM's L1 load-return path: every lane must receive its
                # 8 texels (128 B) per sample through it, whatever the hit rate
                "bound": "l1", "achieved": achieved, "peak": l1_gbs, "unit": "GB/s", "frac": achieved / l1_gbs,
                "traffic": per_view * views_per_launch if per_view else None,
                "traffic_source": f"profiles/{traffic_file} (ncu dram__bytes_read+write of a 1-view launch x views per launch)" if per_view else None,
                "peak_source": "measured live: pyvr_cuda_measure_cache_bandwidth(level 1) -- coalesced LDG.128 hitting L1, "
                               "bytes delivered to registers / CUDA-event time (csrc/bandwidth.cu)",
                "peak_nominal": l1_nominal, "frac_of_nominal": achieved / l1_nominal,
                "kernel": f"march_kernel<fast, {args.texels}x4, {texel_layout}>", "kernel_ms_per_launch": launch_ms, "views_per_launch": views_per_launch,
                "algorithmic_bytes_per_sample": bytes_per_sample,
                "samples_fetched_per_launch": fetched_per_launch,
                "samples_reference_per_launch": samples / max(launches, 1),
                "fetched_Gsamples_per_s": fetched_per_launch / (launch_ms * 1e-3) / 1e9,
                "kernel_share_of_step": kernel_ms / ms if world == 1 else None,
                "l2": {"achieved": achieved, "peak": l2_gbs, "unit": "GB/s", "frac": achieved / l2_gbs,
                       "peak_source": "measured live: pyvr_cuda_measure_cache_bandwidth(level 2) -- coalesced LDG.128.cg over 64 MiB",
                       "note": "SURVEY.md section 8(d)'s L2 form: algorithmic gather bytes against the L2->SM read bandwidth; L1 absorbs "
                               "the 8x gather amplification, so this fraction may exceed 1 and L2 is not the binding unit"},
                "hbm": hbm,
                "dram": ({"achieved": per_view * views_per_launch / (launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": per_view * views_per_launch / (launch_ms * 1e-3) / 1e9 / peak,
                          "note": "DRAM traffic ncu measured per launch / kernel time"} if per_view else None),
                "dense": ({**dense, "achieved": dense["fetched_Gsamples/s"] * bytes_per_sample, "peak": l1_gbs,
                           "frac": dense["fetched_Gsamples/s"] * bytes_per_sample / l1_gbs,
                           "note": "same kernel with empty-space skipping off (every in-box sample fetched), 1 step"} if dense else None),
                "note": "achieved = fetched samples x 8 texels x 16 B / march-kernel time (CUDA events inside the C ABI call, "
                        "on the launching stream); peak = L1 load-return bandwidth measured on this GPU in this run",
            },
            "clocks": clocks.summary(),
            "normals_kernel": normals_info,
            "alternatives": alternatives,
            "setup_s": setup_s,
        }

    # ---------------- N > 1: the partitioned configs and their parity, time-boxed ----------------
    renderer.close()
    pinned.close()
    del d_frames
    torch.cuda.empty_cache()
    if world > 1 and not args.no_secondary:
        import bench_partitioned

        secondary = bench_partitioned.secondary_section(args, rank, world, local_rank, line, args.secondary_seconds)
        if rank == 0:
            line.update(secondary)
    if world == 1 and rank == 0 and not args.no_secondary:
        # N = 1: config C4 (2048^3 f16 scalar+normal, 3840x2160, ultra preset) as ONE frame stream on this GPU, so that the
        # single-GPU number behind the N > 1 tile lines is in the driver's own BENCH line too.  ~15 s, 65 GiB on the device;
        # nothing here may cost the main line.
        try:
            import bench_partitioned

            c4 = bench_partitioned.measure(args, 0, 1, local_rank, "c4", steps=8, warmup=2)
            keep = ("value", "unit", "ms_per_step", "frames_per_s", "samples_per_frame", "config", "e2e", "roofline", "clocks",
                    "volume_generation", "steps", "warmup", "n_gpus")
            line["secondary"] = {"c4": {k: c4[k] for k in keep if k in c4}}
        except Exception as e:
            line["secondary"] = {"c4": {"error": f"{type(e).__name__}: {e}"}}
        torch.cuda.empty_cache()
    if rank == 0 and not args.skip_cpu_baseline and world == 1:
        cores = use_all_host_threads()
        try:
            ref = CpuReference(args, data, normals, light, config, lut, args.reference_backend)
            line["cpu_baseline"] = cpu_baseline(ref, step_view_indices(args.warmup, 0, 1, per_step), args.cpu_seconds)
            line["cpu_baseline"]["host_cores"] = cores
        except Exception as e:      # the GPU numbers above must not be lost to a failure of the CPU arm
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
