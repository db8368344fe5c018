#!/usr/bin/env python
"""Headline benchmark: Gsamples/s of the volume ray march on BASELINE.json's config C3.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (SURVEY.md section 8 d, C3 = BASELINE.json configs[2]): synthetic 512^3 float32
``double_sphere`` volume + normal volume, bounds +-1, 1920x1080, ``RenderConfig.high_quality()``
(step 0.005, 1000 steps, reference step 0.01), ``Light.directional([1,-1,0])``, viridis colour TF +
``linear(0, 0.1)`` opacity TF, the 360-view turntable (azimuth 2*pi*k/360, elevation pi/6, distance 3).

A *step* is one batch of ``--views-per-step`` turntable views per GPU.  Views are dealt ``k mod N``
over the N ranks (no data-path collective; every rank holds the whole volume), so per-GPU work is
fixed as N grows: ``"scaling": "weak"``.  A *sample* is one in-box loop body of the reference shader
(volume.frag.glsl:92-116), counted by the kernel; ``value`` = samples of all ranks / max-over-ranks
device time, frames device-resident.  ``e2e`` = the same through ``VolumeRenderer.render_batch`` with
the views coming from pinned host memory and the RGBA8 frames read back to pinned host memory inside
the timed region.

``--impl reference`` times the CPU restatement of the reference shader (oracle/, C + OpenMP, all host
threads) on a bounded sample of the same workload -- the first few whole views of every step: the
reference itself needs moderngl + an OpenGL driver, which do not exist in this image ("kind": "port").
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
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 = headline turntable (default); c4 = image tiles over a replicated device-generated "
                         "volume; c5 = sort-last bricks + binary swap (secondary lines, see DESIGN.md)")
    ap.add_argument("--size", type=int, default=None, help="volume edge (default 512 / 2048 / 2048*cbrt(N))")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="c5: binary-swap exchange path")
    ap.add_argument("--views-per-step", type=int, default=12)
    ap.add_argument("--texels", default="f32", choices=["f32", "f16"])
    ap.add_argument("--no-ess", action="store_true", help="disable empty-space skipping")
    ap.add_argument("--hwtex", action="store_true", help="sample through the texture unit (hardware trilinear)")
    ap.add_argument("--layout", default=None, choices=["linear", "swizzle"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-alternatives", action="store_true", help="skip the f16 / hwtex side measurements")
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


def cpu_baseline(data, normals, light, config, lut, args, view_indices, target_seconds):
    """Oracle (C + OpenMP restatement of the reference shader) on a bounded sample of the same workload: whole
    views of the step, one after the other, until `target_seconds` of CPU time have been spent."""
    import oracle
    from pyvr_b200 import Volume

    vol = Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32),
                 max_bounds=np.array([1, 1, 1], np.float32))
    samples, views, t0 = 0, 0, time.perf_counter()
    for k in view_indices:
        _, _, st = oracle.render(vol, turntable_camera(k), light, config, lut, args.width, args.height)
        samples += st["samples"]
        views += 1
        if time.perf_counter() - t0 >= target_seconds:
            break
    dt = time.perf_counter() - t0
    return {
        "value": samples / dt / 1e9, "unit": "Gsamples/s", "cores": oracle.num_threads(), "kind": "port",
        "sample": f"{views} whole views of the step (turntable views {view_indices[0]}..{view_indices[views - 1]}, "
                  f"{samples} samples) in {dt:.2f} s; CPU restatement of the reference shader "
                  f"(oracle/pyvr_oracle.c, OpenMP), llvmpipe/moderngl unavailable in this image",
        "seconds": dt, "frames_per_s": views / dt,
    }


def run_reference(args, rank):
    """--impl reference: the CPU path on rank 0 only."""
    if rank != 0:
        return
    import oracle
    from pyvr_b200 import Volume

    data, light, config, lut = scene(args.size)
    normals = oracle.normals(data)   # the reference's compute_normal_volume, restated (bit-exact vs numpy)
    vol = Volume(data=data, normals=normals, min_bounds=np.array([-1, -1, -1], np.float32),
                 max_bounds=np.array([1, 1, 1], np.float32))
    per_step = args.views_per_step
    # bounded sample: the first `n_sample` WHOLE views of each step's 12 (whole views keep all host threads busy;
    # a few rows per view would not), sized from one calibration view for about 3 s of CPU time per step
    t0 = time.perf_counter()
    oracle.render(vol, turntable_camera(0), light, config, lut, args.width, args.height)
    t_view = time.perf_counter() - t0
    n_sample = int(min(per_step, max(1, round(3.0 / max(t_view, 1e-6)))))

    def step(s):
        total = 0
        for k in step_view_indices(s, 0, 1, per_step)[:n_sample]:
            _, _, st = oracle.render(vol, turntable_camera(k), light, config, lut, args.width, args.height)
            total += st["samples"]
        return total

    for s in range(args.warmup):
        step(s)
    t0 = time.perf_counter()
    samples = sum(step(args.warmup + s) for s in range(args.steps))
    dt = time.perf_counter() - t0
    value = samples / dt / 1e9
    sample = (f"the first {n_sample} whole views of each step's {per_step}; CPU restatement of the reference shader "
              "(oracle/, C + OpenMP); the reference's ModernGL path cannot run here (no moderngl, no OpenGL driver)")
    line = {
        "impl": "reference", "metric": "ray-march throughput", "value": value, "unit": "Gsamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "frames_per_s": args.steps * n_sample / dt,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, stride_note=f"{n_sample} of {per_step} views per step"),
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": oracle.num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, stride_note=None):
    cfg = {
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
    if stride_note:
        cfg["sample"] = stride_note
    return cfg


def time_normals_kernel(torch, _cabi, data, device):
    """K2 on device-resident buffers: best of 5 launches after a warm-up (CUDA events inside the C ABI call)."""
    n0, n1, n2 = data.shape
    d_in = torch.from_numpy(data).cuda(device)
    d_out = torch.empty((n0, n1, n2, 3), dtype=torch.float32, device=d_in.device)
    ms, best = ctypes.c_float(0.0), float("inf")
    for _ in range(6):
        _cabi.check(_cabi.lib().pyvr_cuda_compute_normals(device, ctypes.c_void_p(d_in.data_ptr()),
                                                          ctypes.c_void_p(d_out.data_ptr()), n0, n1, n2, 1, ctypes.byref(ms)))
        best = min(best, ms.value)
    del d_in, d_out
    torch.cuda.empty_cache()
    return best


def traffic_per_view(args):
    """DRAM bytes one view of the march moves, from the committed ncu capture (profiles/r01_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        table = json.load(open(path))
        key = f"{args.texels}_{'hwtex_' if args.hwtex else ''}{'dense' if args.no_ess else 'ess'}"
        return float(table["c3_dram_bytes_per_view"][key])
    except Exception:
        return None


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

    t_setup = time.perf_counter()
    renderer_kw = dict(device=local_rank, texel_format=args.texels, empty_space_skipping=not args.no_ess,
                       hardware_filtering=args.hwtex)
    if world == 1:
        # host pipeline, as a user of the reference would: create_sample_volume -> compute_normal_volume (K2 on
        # the GPU) -> Volume -> load_volume.  The host arrays also feed the CPU baseline.
        data, light, config, lut = scene(args.size)
        normals, _ = _cabi.compute_normals_host(data, device=local_rank, return_ms=True)   # K2 (first launch: cold)
        normals_ms = time_normals_kernel(torch, _cabi, data, local_rank)
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
        normals_ms = renderer.generate_volume(args.size, "double_sphere", (-1, -1, -1), (1, 1, 1))
    renderer.set_lut(lut)
    stream = torch.cuda.Stream()   # non-default: the library treats stream 0 as "use the context's own stream"
    renderer.set_stream(stream.cuda_stream)
    setup_s = time.perf_counter() - t_setup

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
    def resident_step(s):
        renderer.render_batch(views=views[s], device_ptr=d_frames.data_ptr())
        return renderer.stats

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

    # ---------------- alternatives (N = 1 only): same views, other texel storage / sampler, device-resident.
    # Not the headline: the headline is binary32 texels + binary32 software trilinear (the north star's
    # primary path); these are the configurations the north star lists as allowed when within tolerance.
    alternatives = {}
    if world == 1 and not args.no_alternatives and not args.hwtex and args.texels == "f32":
        for name, kw in (("f16x4 texels, software trilinear", dict(texel_format="f16")),
                         ("f16x4 texels, texture-unit trilinear (hwtex)", dict(texel_format="f16", hardware_filtering=True))):
            alt = VolumeRenderer(args.width, args.height, config=config, light=light, device=local_rank,
                                 empty_space_skipping=not args.no_ess, **kw)
            alt.load_volume(vol)
            alt.set_lut(lut)
            alt.set_stream(stream.cuda_stream)
            for s_ in range(args.warmup):
                alt.render_batch(views=views[s_], device_ptr=d_frames.data_ptr())
            torch.cuda.synchronize()
            e0.record(stream)
            alt_samples = 0
            for s_ in range(args.warmup, total_steps):
                alt.render_batch(views=views[s_], device_ptr=d_frames.data_ptr())
                alt_samples += alt.stats["samples"]
            e1.record(stream)
            torch.cuda.synchronize()
            alt_ms = e0.elapsed_time(e1)
            alternatives[name] = {"value": alt_samples / (alt_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                                  "frames_per_s": args.steps * per_step / (alt_ms * 1e-3)}
            alt.close()

    if rank == 0:
        bytes_per_sample = BYTES_PER_SAMPLE_F32 if args.texels == "f32" else BYTES_PER_SAMPLE_F16
        peak, peak_src = measured_peak_gbs()
        launch_ms = kernel_ms / max(launches, 1)
        achieved = (fetched / max(launches, 1)) * bytes_per_sample / (launch_ms * 1e-3) / 1e9
        views_per_launch = per_step if per_step <= 16 else 16
        per_view = traffic_per_view(args)
        sm_mhz = clocks.summary().get("sm_mhz") or 1965.0
        l1_peak = 148 * 128 * sm_mhz * 1e6 / 1e9      # one 128-byte L1 wavefront per SM per clock
        line = {
            "metric": "ray-march throughput", "value": value, "unit": "Gsamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "frames_per_s": frames / (ms * 1e-3),
            "samples_per_frame": all_samples / frames,
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "frames_per_s": frames / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": world * per_step * ctypes.sizeof(_cabi.View),
                    "d2h_bytes_per_step": world * per_step * frame_bytes,
                    "api": "VolumeRenderer.render_batch(views, out=pinned host buffer)", "frame_checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": per_view * views_per_launch if per_view else None,
                "traffic_source": "profiles/r01_traffic.json (ncu dram__bytes_read+write of a 1-view launch x views per launch)",
                "peak_source": peak_src, "kernel": "march_kernel<fast>",
                "kernel_ms_per_launch": launch_ms, "views_per_launch": views_per_launch,
                "l1": {"bound": "l1tex data stage", "achieved": achieved, "peak": l1_peak, "unit": "GB/s",
                       "frac": achieved / l1_peak,
                       "note": "same algorithmic bytes against 148 SMs x 128 B/clk at the sampled SM clock: the unit "
                               "that actually binds this gather kernel (ncu: l1tex data-stage 71-76 % busy, issue slots 58-68 %, DRAM 16-24 %)"},
                "dram": ({"achieved": per_view * views_per_launch / (launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": per_view * views_per_launch / (launch_ms * 1e-3) / 1e9 / peak,
                          "note": "measured DRAM traffic per launch / kernel time: each packed line comes from HBM about "
                                  "once per view, so HBM is far from binding"} if per_view else None),
                "algorithmic_bytes_per_sample": bytes_per_sample,
                "samples_fetched_per_launch": fetched / max(launches, 1),
                "samples_reference_per_launch": samples / max(launches, 1),
                "kernel_share_of_step": kernel_ms / ms if world == 1 else None,
                "note": "achieved = fetched samples x 8 texels x texel bytes / march-kernel time.  The gather is served "
                        "by L1/L2 (each packed line is read from HBM about once per view: `traffic`), so the "
                        "fraction of the HBM copy rate exceeds 1; see `l1` for the binding unit",
            },
            "clocks": clocks.summary(),
            "normals_kernel": ({"ms": normals_ms, "GB/s": args.size ** 3 * 16 / (normals_ms * 1e-3) / 1e9,
                                "frac_of_hbm_peak": args.size ** 3 * 16 / (normals_ms * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes_per_voxel": 16} if world == 1 else
                               {"device_generation_ms": normals_ms}),
            "alternatives": alternatives,
            "setup_s": setup_s,
        }
        if not args.skip_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(data, normals, light, config, lut, args,
                                                step_view_indices(args.warmup, 0, 1, per_step), args.cpu_seconds)
        print(json.dumps(line), flush=True)

    renderer.close()
    pinned.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
