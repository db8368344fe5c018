"""Secondary bench lines for the partitioned configs of BASELINE.json (run through ``bench.py --workload``).

    python bench.py --workload c4 [--size 2048]                       # 1 GPU
    torchrun --nproc-per-node N bench.py --workload c4 --gpus N       # image tiles over N GPUs
    torchrun --nproc-per-node 8 bench.py --workload c5 --gpus 8       # 4096^3 sort-last + binary swap

* **c4** (configs[3]): synthetic ``size``^3 (default 2048) f16 scalar+normal volume generated on the device
  and replicated on every GPU, 3840x2160, ``ultra_quality``, one isometric view, empty-space skipping;
  64x64-pixel tile groups are dealt round-robin over the ranks and one ``reduce(SUM)`` of the uint8 frames
  (NCCL) assembles the frame on rank 0.  A step = one frame; ``"scaling": "strong"`` (the frame is fixed).
* **c5** (configs[4]): synthetic ``size``^3 f16 volume (default: 2048 * N^(1/3), i.e. 4096^3 on 8 GPUs,
  one 2048^3(+ghost) brick = 68.8 GB per GPU) generated brick by brick on the device, 3840x2160,
  ``high_quality``; every rank marches its brick into a float4 partial image, binary swap over NVLink
  (``--exchange p2p``: the merge kernel reads the partner's half through a CUDA-IPC mapping; ``nccl``:
  send/recv + local merge), finalise + gather on rank 0.  A step = one frame; ``"scaling": "weak"``
  (per-GPU voxels fixed).

``value`` = reference samples of all ranks / max-over-ranks device time of the whole step (march +
exchange + merge + gather), frames device-resident; ``e2e`` adds the camera upload and the read-back of the
assembled RGBA8 frame into pinned host memory on rank 0.
"""

from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import bench
    from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                           build_rgba_lut)
    from pyvr_b200 import multi_gpu as mg
    from pyvr_b200.cuda_renderer import VolumeRenderer, _cabi

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the partitioned configs (c4/c5) have no CPU arm: "
                              "2048^3+ volumes are generated on the device; use the default workload"}), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyvr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    c5 = args.workload == "c5"
    if args.size is None:
        args.size = int(round(2048 * round(world ** (1 / 3)))) if c5 else 2048
    width, height = args.width or 3840, args.height or 2160
    n_pixels = width * height
    config = RenderConfig.high_quality() if c5 else RenderConfig.ultra_quality()
    light = Light.directional([1, -1, 0])
    lut = build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.1))
    camera = Camera.isometric_view(distance=3.0)
    position, _ = camera.get_camera_vectors()
    bmin, bmax = (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)
    texels = "f16" if args.texels == "f32" and args.size >= 1024 else args.texels   # north-star storage for C4/C5

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    renderer = VolumeRenderer(width, height, config=config, light=light, device=local_rank, texel_format=texels,
                              empty_space_skipping=not args.no_ess, hardware_filtering=args.hwtex)
    renderer.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    shape = (args.size,) * 3
    if c5 and world > 1:
        brick = mg.brick_of_rank(shape, rank, world)
        gen_ms = renderer.generate_volume(args.size, "double_sphere", bmin, bmax, brick=brick)
        session = mg.SortLastSession(shape, bmin, bmax, n_pixels, device=local_rank, exchange=args.exchange)
        stored = brick.dims
    else:
        brick, session = None, None
        gen_ms = renderer.generate_volume(args.size, "double_sphere", bmin, bmax)
        stored = shape
        if world > 1:
            renderer.set_pixel_shard(rank, world)
    renderer.set_lut(lut)
    renderer.set_camera(camera)
    setup_s = time.perf_counter() - t0

    frame = torch.zeros((n_pixels, 4), dtype=torch.uint8, device="cuda")
    frame_ptr = frame.data_ptr()
    pinned = _cabi.PinnedBuffer(n_pixels * 4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def step(read_back=False):
        """One frame.  Returns this rank's pyvr_stats."""
        nonlocal frame, frame_ptr
        if read_back:
            renderer.set_camera(camera)                      # camera uniforms travel host -> device again
        if session is not None:
            renderer.render_accum_to_device(session.image_ptr())
            st = renderer.stats
            piece_range, piece = session.composite(position)
            out = session.gather_rgba8(piece_range, piece)
            if out is not None:
                frame_ptr = out if isinstance(out, int) else out.data_ptr()
        else:
            renderer.render_to_device(frame.data_ptr())
            st = renderer.stats
            if world > 1:
                mg.reduce_tile_frames(frame, dst=0)
        if read_back and rank == 0:
            _cabi.check(_cabi.lib().pyvr_cuda_memcpy(local_rank, pinned.array.ctypes.data, ctypes.c_void_p(frame_ptr),
                                                     n_pixels * 4, 2, ctypes.c_void_p(stream.cuda_stream)))
        return st

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = fetched = 0
    kernel_ms = 0.0
    with bench.ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            st = step()
            samples += st["samples"]
            fetched += st["samples_fetched"]
            kernel_ms += st["kernel_ms"]
        e1.record(stream)
        barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    all_samples, all_fetched = reduce_sum(float(samples)), reduce_sum(float(fetched))
    march_ms = reduce_max(kernel_ms)

    for _ in range(args.warmup):
        step(read_back=True)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step(read_back=True)
    e1.record(stream)
    barrier()
    e2e_ms = reduce_max(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    nonzero = int(np.count_nonzero(pinned.array[3::4])) if rank == 0 else 0

    if rank == 0:
        bytes_per_sample = 64 if texels == "f16" else 128
        peak, peak_src = bench.measured_peak_gbs()
        launch_ms = march_ms / args.steps
        achieved = (all_fetched / world / args.steps) * bytes_per_sample / (launch_ms * 1e-3) / 1e9
        sm_mhz = clocks.summary().get("sm_mhz") or 1965.0
        l1_peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        voxels = float(np.prod(stored))
        name = (f"C5: synthetic {args.size}^3 f16 scalar+normal double_sphere generated on device, {world} sort-last bricks "
                f"of {stored[0]}x{stored[1]}x{stored[2]} (+1 ghost), {width}x{height}, high_quality, binary swap ({args.exchange}) "
                "over NVLink, isometric view" if c5 else
                f"C4: synthetic {args.size}^3 f16 scalar+normal double_sphere generated on device and replicated, {width}x{height}, "
                f"ultra_quality, ESS, isometric view, 64x64 tile groups round-robin over {world} GPU(s), reduce(SUM) of uint8 frames")
        line = {
            "metric": "ray-march throughput", "value": all_samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if c5 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "texels": "f16x4 (8 B/voxel)" if texels == "f16" else "f32x4 (16 B/voxel)",
                       "l2_policy": f"inputs larger than L2 (packed block {voxels * (8 if texels == 'f16' else 16) / 2 ** 30:.1f} GiB per GPU), no flush",
                       "empty_space_skipping": not args.no_ess,
                       "sampling": "texture unit, hardware trilinear (8-bit weights)" if args.hwtex else "binary32 software trilinear"},
            "frames_per_s": args.steps / (ms * 1e-3), "samples_per_frame": all_samples / args.steps,
            "e2e": {"value": all_samples / (e2e_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                    "frames_per_s": args.steps / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": world * ctypes.sizeof(_cabi.View), "d2h_bytes_per_step": n_pixels * 4,
                    "api": "VolumeRenderer.set_camera + render (+ multi_gpu exchange) + frame read-back to pinned host memory",
                    "frame_nonzero_alpha_pixels": nonzero},
            "gpu_launches": int(args.steps * world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "march_kernel<fast>",
                         "kernel_ms_per_launch": launch_ms, "algorithmic_bytes_per_sample": bytes_per_sample,
                         "march_share_of_step": march_ms / ms,
                         "l1": {"achieved": achieved, "peak": l1_peak, "unit": "GB/s", "frac": achieved / l1_peak},
                         "note": "per-GPU figures of the slowest rank's march; exchange/merge/gather are the rest of the step"},
            "clocks": clocks.summary(),
            "volume_generation": {"ms": gen_ms, "voxels_per_gpu": voxels, "Gvoxels/s": voxels / (gen_ms * 1e-3) / 1e9},
            "setup_s": setup_s,
        }
        print(json.dumps(line), flush=True)

    if session is not None:
        session.close()
    renderer.close()
    pinned.close()
    if world > 1:
        dist.destroy_process_group()
