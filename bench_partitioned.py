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

import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def default_c5_size(world):
    """Edge of the C5 volume for `world` bricks: per-GPU voxels stay at 2048^3 (weak scaling), multiple of 64."""
    return int(round(2048 * world ** (1 / 3) / 64) * 64)


def run(args, rank, world, local_rank):
    """``bench.py --workload c4|c5``: one JSON line for the partitioned config."""
    import torch
    import torch.distributed as dist

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
    line = measure(args, rank, world, local_rank, args.workload, args.steps, args.warmup, args.size, args.width, args.height)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure(args, rank, world, local_rank, workload, steps, warmup, size=None, width=None, height=None):
    """Measure one partitioned config on an initialised process group.  Returns the JSON-able line on rank 0,
    None elsewhere."""
    import torch
    import torch.distributed as dist

    import bench
    from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig,
                           build_rgba_lut)
    from pyvr_b200 import multi_gpu as mg
    from pyvr_b200.cuda_renderer import VolumeRenderer, _cabi

    args = argparse.Namespace(**vars(args))
    args.workload, args.steps, args.warmup = workload, steps, warmup
    args.size, args.width, args.height = size, width, height     # None = the config's own (2048 / 2048*cbrt(N), 3840x2160)
    c5 = args.workload == "c5"
    if args.size is None:
        args.size = default_c5_size(world) if c5 else 2048
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
        session = mg.SortLastSession(shape, bmin, bmax, n_pixels, device=local_rank, exchange=args.exchange, renderer=renderer)
        stored = brick.dims
    else:
        brick, session = None, None
        gen_ms = renderer.generate_volume(args.size, "double_sphere", bmin, bmax)
        stored = shape
    tiles = mg.TileSession(renderer, n_pixels, device=local_rank) if (not c5 and world > 1) else None
    renderer.set_lut(lut)
    renderer.set_camera(camera)
    setup_s = time.perf_counter() - t0

    texel_layout = renderer.texel_layout
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
        # the sessions make device-output renders asynchronous: the exchange is enqueued behind the march while it
        # runs, and the counters are read (renderer.stats waits for them) only after everything has been enqueued
        if session is not None:
            renderer.render_accum_to_device(session.image_ptr())
            piece_range, piece = session.composite(position, finalize_to=0 if args.exchange == "p2p" else None)
            out = session.gather_rgba8(piece_range, piece)
            if out is not None:
                frame_ptr = out if isinstance(out, int) else out.data_ptr()
        elif tiles is not None:
            out = tiles.render()                 # march + peer stores into rank 0's frame + two flags
            if out is not None:
                frame_ptr = out
        else:
            renderer.render_to_device(frame.data_ptr())
        if read_back and rank == 0:
            _cabi.check(_cabi.lib().pyvr_cuda_memcpy(local_rank, pinned.array.ctypes.data, ctypes.c_void_p(frame_ptr),
                                                     n_pixels * 4, 2, ctypes.c_void_p(stream.cuda_stream)))
        if tiles is not None:
            tiles.release()
        return renderer.stats

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
    march_ms_mean = reduce_sum(kernel_ms) / world

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
        # DRAM bytes one whole C4 frame moves (committed ncu capture of the same layout), for the 1-GPU roofline
        traffic = dram = None
        if not c5 and world == 1 and not args.hwtex:
            try:
                table = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")))
                key = ("bricks_four_in_flight" if "several" in texel_layout else "bricks_one_sample") if "bricks" in texel_layout \
                    else "rows_zpairs_one_sample"
                traffic = float(table["c4_dram_bytes_per_frame"][key])
                dram = {"achieved": traffic / (launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": traffic / (launch_ms * 1e-3) / 1e9 / peak,
                        "note": f"DRAM traffic ncu measured for one frame of this layout ({key}, profiles/r02_traffic.json) / kernel time: "
                                "the binding unit of the sparse-ray march"}
            except Exception:
                traffic = dram = None
        name = (f"C5: synthetic {args.size}^3 f16 scalar+normal double_sphere generated on device, {world} sort-last bricks "
                f"of {stored[0]}x{stored[1]}x{stored[2]} (+1 ghost), {width}x{height}, high_quality, binary swap ({args.exchange}) "
                "over NVLink, isometric view" if c5 else
                f"C4: synthetic {args.size}^3 f16 scalar+normal double_sphere generated on device and replicated, {width}x{height}, "
                f"ultra_quality, ESS, isometric view, 32x16-pixel tile groups dealt over {world} GPU(s), pixels stored straight into "
                "rank 0's frame over NVLink by the march kernel (no reduce / gather pass)")
        line = {
            "metric": "ray-march throughput", "value": all_samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if c5 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "texels": "f16x4 (8 B/voxel)" if texels == "f16" else "f32x4 (16 B/voxel)",
                       "texel_layout": texel_layout,
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
                         "traffic": traffic, "dram": dram, "peak_source": peak_src, "kernel": "march_kernel<fast>",
                         "kernel_ms_per_launch": launch_ms, "algorithmic_bytes_per_sample": bytes_per_sample,
                         "march_share_of_step": march_ms / ms,
                         "l1": {"achieved": achieved, "peak": l1_peak, "unit": "GB/s", "frac": achieved / l1_peak},
                         "note": "per-GPU figures of the slowest rank's march; exchange/merge/gather are the rest of the step"},
            "balance": {"march_ms_slowest_rank": march_ms / args.steps, "march_ms_mean_over_ranks": march_ms_mean / args.steps,
                        "imbalance": march_ms / march_ms_mean if march_ms_mean > 0 else None},
            "clocks": clocks.summary(),
            "volume_generation": {"ms": gen_ms, "voxels_per_gpu": voxels, "Gvoxels/s": voxels / (gen_ms * 1e-3) / 1e9},
            "setup_s": setup_s,
        }
    else:
        line = None

    if session is not None:
        session.close()
    if tiles is not None:
        tiles.close()
    renderer.close()
    pinned.close()
    del frame
    torch.cuda.empty_cache()
    return line


def parity_check(rank, world, local_rank):
    """The multi-process paths against the single-GPU frame, on C1's volume (128^3 double_sphere + normals, balanced,
    isometric view) at 400x300: (1) image tiles: reduce(SUM) of the per-rank frames == the single-GPU frame, bit for
    bit; (2) sort-last bricks + binary swap (p2p and nccl exchange): within max |delta| <= 3/255, >= 99.9 % of the
    pixels within 1/255 (the merge clips saturating rays to alpha 0.99); (3) relay: bit-identical; (4) on rank 0 the single-GPU frame itself against the
    CPU oracle within BASELINE.json's tolerance.  Returns a dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig, Volume,
                           build_rgba_lut, compute_normal_volume, create_sample_volume)
    from pyvr_b200 import multi_gpu as mg
    from pyvr_b200.cuda_renderer import VolumeRenderer, _cabi

    W, H = 400, 300
    data = create_sample_volume(128, "double_sphere")
    vol = Volume(data=data, normals=compute_normal_volume(data))
    light, cfg = Light.directional([1, -1, 0]), RenderConfig.balanced()
    lut = build_rgba_lut(ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.3))
    cam = Camera.isometric_view(distance=3.0)
    position, _ = cam.get_camera_vectors()
    out = {"scene": "C1 volume (128^3 double_sphere + normals), balanced, isometric view, 400x300", "ranks": world}

    def frame_of(ptr_or_tensor):
        if isinstance(ptr_or_tensor, int):
            torch.cuda.synchronize()
            host = np.empty(W * H * 4, np.uint8)
            _cabi.check(_cabi.lib().pyvr_cuda_memcpy(local_rank, host.ctypes.data, ctypes.c_void_p(ptr_or_tensor), W * H * 4, 2, None))
            return host.reshape(H, W, 4)
        return ptr_or_tensor.cpu().numpy().reshape(H, W, 4)

    def metrics(got, want):
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        return {"max_abs": int(d.max()), "frac_within_1": float((d <= 1).all(axis=-1).mean()),
                "frac_identical": float((d == 0).all(axis=-1).mean())}

    with VolumeRenderer(W, H, config=cfg, light=light, device=local_rank) as r:
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        r.load_volume(vol)
        r.set_camera(cam)
        r.set_lut(lut)
        want = np.frombuffer(r.render(), np.uint8).reshape(H, W, 4).copy()
        want_samples = r.stats["samples"]
        # (1) image tiles: cleared frames + reduce(SUM), and the fused path (peer stores into rank 0's frame)
        r.set_pixel_shard(rank, world)
        tiles = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
        r.render_to_device(tiles.data_ptr())
        mg.reduce_tile_frames(tiles, dst=0)
        r.set_pixel_shard(0, 1)
        ts = mg.TileSession(r, W * H, device=local_rank)
        for _ in range(2):
            ptr = ts.render()
            if ptr is not None:
                fused = frame_of(ptr).copy()
            ts.release()
        ts.close()
        if rank == 0:
            out["tiles_bit_identical"] = bool(np.array_equal(frame_of(tiles), want))
            out["tiles_fused_bit_identical"] = bool(np.array_equal(fused, want))
        # (2) sort-last bricks, both exchange paths; (3) relay
        if world & (world - 1) == 0:
            b = mg.brick_of_rank(data.shape, rank, world)
            r.load_brick(vol.data[b.slices()], vol.normals[b.slices()], data.shape, b.origin, b.own_lo, b.own_hi,
                         vol.min_bounds, vol.max_bounds)
            r.set_camera(cam)
            r.set_lut(lut)
            for exchange in ("p2p", "nccl"):
                session = mg.SortLastSession(data.shape, vol.min_bounds, vol.max_bounds, W * H, device=local_rank, exchange=exchange, renderer=r)
                for it in range(3):                  # several frames: buffers and flags are reused; the last one fused
                    r.render_accum_to_device(session.image_ptr())
                    piece_range, piece = session.composite(position, finalize_to=0 if (exchange == "p2p" and it == 2) else None)
                    frame = session.gather_rgba8(piece_range, piece)
                samples = torch.tensor([r.stats["samples"]], dtype=torch.int64, device="cuda")
                dist.all_reduce(samples)
                if rank == 0:
                    m = metrics(frame_of(frame), want)
                    m["ok"] = bool(m["max_abs"] <= 3 and m["frac_within_1"] >= 0.999)
                    # every sample lands in exactly one brick, but a back brick cannot see that the shader stopped the
                    # ray in front of it (alpha >= 0.99), so the bricks execute a few more samples than the single pass
                    m["brick_samples_over_single_gpu"] = int(samples.item()) / max(want_samples, 1)
                    out[f"sort_last_{exchange}"] = m
                session.close()
            relay = mg.RelaySession(data.shape, vol.min_bounds, vol.max_bounds, W * H, device=local_rank)
            relay_frame = relay.render(r, position)
            flag = torch.zeros(1, dtype=torch.int32, device="cuda")
            if relay_frame is not None:              # the last rank of the visibility order holds the frame
                flag[0] = int(np.array_equal(frame_of(relay_frame), want)) + 1
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if rank == 0:
                out["relay_bit_identical"] = bool(int(flag.item()) == 2)
    if rank != 0:
        return None
    try:   # the checker: single-GPU frame vs the CPU oracle (BASELINE tolerance)
        import oracle

        ref, _, _ = oracle.render(vol, cam, light, cfg, lut, W, H)
        d = np.abs(want.astype(np.int32) - ref.astype(np.int32))
        mse = float(np.mean((want.astype(np.float64) - ref.astype(np.float64)) ** 2))
        out["single_gpu_vs_oracle"] = {"max_abs": int(d.max()), "frac_within_2": float((d <= 2).all(axis=-1).mean()),
                                       "psnr_db": float("inf") if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))}
    except Exception as e:
        out["single_gpu_vs_oracle"] = {"error": f"{type(e).__name__}: {e}"}
    checks = [out.get("tiles_bit_identical", False), out.get("tiles_fused_bit_identical", False)]
    if world & (world - 1) == 0:
        checks += [out["sort_last_p2p"]["ok"], out["sort_last_nccl"]["ok"], out["relay_bit_identical"]]
    out["all_ok"] = bool(all(checks))
    return out


def secondary_section(args, rank, world, local_rank, line, seconds):
    """N > 1: C4 (image tiles) and C5 (sort-last) lines plus the parity check, inside a time box.  A watchdog
    thread makes sure the main line is still printed (and every rank exits 0) if a collective of this section
    hangs.  Returns {"secondary": {...}, "parity_check": {...}} on rank 0."""
    import threading

    result = {"secondary": {}, "parity_check": None}

    def give_up():
        if rank == 0:
            result["secondary"]["error"] = f"secondary section exceeded its {seconds:.0f} s time box; partial results kept"
            line.update(result)
            print(json.dumps(line), flush=True)
        os._exit(0)

    dog = threading.Timer(seconds, give_up)
    dog.daemon = True
    dog.start()
    t0 = time.perf_counter()
    try:
        try:
            result["parity_check"] = parity_check(rank, world, local_rank)
        except Exception as e:      # a failure here is deterministic (same on every rank): report and go on
            result["parity_check"] = {"error": f"{type(e).__name__}: {e}"}
        for workload in ("c4", "c5"):
            if workload == "c5" and world & (world - 1):
                continue
            try:
                got = measure(args, rank, world, local_rank, workload, steps=8, warmup=2)
                if rank == 0:
                    keep = ("value", "unit", "ms_per_step", "frames_per_s", "samples_per_frame", "scaling", "config", "e2e",
                            "roofline", "clocks", "volume_generation", "steps", "warmup", "n_gpus", "balance")
                    result["secondary"][workload] = {k: got[k] for k in keep if k in got}
            except Exception as e:
                result["secondary"][workload] = {"error": f"{type(e).__name__}: {e}"}
        result["secondary"]["seconds"] = time.perf_counter() - t0
    finally:
        dog.cancel()
    return result if rank == 0 else None
