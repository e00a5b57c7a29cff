#!/usr/bin/env python3
"""bench.py -- Mpixel/s of the compute render path on Ghostscript_Tiger at 8192x8192 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one frame: the encoded scene (resident in HBM) -> binning kernels (k_seg, k_row) -> fill/blend
(k_heavy: the tiles with more than 16 records, one CTA or warp each; k_fine: every other tile, and the framebuffer)
-> RGBA8 framebuffer in HBM.  With N GPUs the frame's tile rows are sharded into N contiguous row-strips (strong
scaling: the frame is fixed); the scene is broadcast once over NCCL before the timed region and there is no
collective per frame.  Rank 0 prints ONE JSON line.

  value      whole-frame Mpixel/s: K frames, each bracketed by a CUDA event pair on the render stream; the step time
             is the mean per-frame time, max over ranks.  The same method for every N; when a rank's strip fits the
             L2, every rank flushes the L2 between frames (outside the event pairs)
  e2e        the same metric through the C-ABI call pm_renderer_render_host: scene bytes in pinned host memory ->
             H2D -> validate + plan -> frame -> D2H of the strip's pixels into pinned host memory
  roofline   fill/blend kernel k_fine: algorithmic bytes (4*W*H_strip - heavy tiles + scene) / its CUDA-event
             duration (second pass, every kernel group timed on its own), against the measured HBM copy bandwidth
             in MEASURED_PEAKS.json
  cpu_baseline   the oracle (CPU port of the reference's tile loop, pinned to the reference's own shader) on the
             host cores: whole frames of the same workload for about 10 s

--impl reference times that CPU port alone on whole frames of the same workload, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "mpixel_per_s_tiger_8192"
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="surface edge in pixels (BASELINE: 8192)")
    ap.add_argument("--scene", default="tiger", choices=["tiger", "rand_bezier", "glyphs"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frame-events", action="store_true", help="(kept for old scripts; the kernel-time pass always runs)")
    ap.add_argument("--equal-strips", action="store_true", help="equal-height row strips instead of cost-balanced ones")
    ap.add_argument("--emulate-world", default="", help="experiments on 1 GPU: 'N:g' renders the strip rank g of N would own")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled through NVML by a thread during the timed
    region (nvidia-smi -lms is too slow to start for a region of tens of milliseconds)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.sm, self.bits, self.smax, self._stop, self._thr = index, [], 0, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[index]) if index < len(ids) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.nv is None:
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop = True
        if self._thr:
            self._thr.join(timeout=1.0)
        reasons = sorted(n for n, b in self.REASONS.items() if self.bits & b)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(self.sm)}


def scene_for(pm, args):
    kind = {"tiger": pm.SCENE_TIGER, "rand_bezier": pm.SCENE_RAND_BEZIER, "glyphs": pm.SCENE_GLYPHS}[args.scene]
    return pm.build_scene(kind, args.size, args.size)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(args):
    return "Ghostscript_Tiger %dx%d" % (args.size, args.size) if args.scene == "tiger" else "%s %dx%d" % (args.scene, args.size, args.size)


def cpu_frames(scene, size, threads, n_frames=None, min_seconds=None):
    """The oracle (all host cores) on WHOLE frames of the workload.  Either exactly n_frames, or as many as
    fill min_seconds.  Returns (frames, seconds)."""
    import oracle_api
    nty = (size + 15) // 16
    oracle_api.render(scene, size, size, tile_y0=0, tile_y1=min(2, nty), threads=threads)  # warm the thread pool
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle_api.render(scene, size, size, threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if (n_frames is not None and reps >= n_frames) or (min_seconds is not None and dt >= min_seconds):
            return reps, dt


def run_reference(args, rank):
    """--impl reference: the reference's tile loop on this box's host cores -- oracle/pm_oracle.c, the CPU port that
    tests/test_ref_pin.py pins bit for bit to the reference's own PietRender.metal (oracle/_ref; that build is
    limited to the reference's 4096 x 4096 surfaces and is 6x slower than the port, so the port is what is timed).
    Every step renders the WHOLE frame of the same workload; if steps + warmup full frames would take more than
    ~4 minutes, the step count is cut (and reported), never the frame."""
    if rank != 0:
        return
    import __graft_entry__ as ge
    pm = ge.load_package()
    scene = scene_for(pm, args)
    threads = host_threads()
    _, t1 = cpu_frames(scene, args.size, threads, n_frames=1)
    budget = 240.0
    warmup = max(0, min(args.warmup, int(0.1 * budget / t1)))
    steps = max(1, min(args.steps, int(0.9 * budget / t1)))
    if warmup:
        cpu_frames(scene, args.size, threads, n_frames=warmup)
    _, dt = cpu_frames(scene, args.size, threads, n_frames=steps)
    value = args.size * args.size * steps / dt / 1e6
    sample = "%d whole frames, %.1f s wall on %d threads" % (steps, dt, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "tile": "16x16",
                   "implementation": "CPU port of the PietRender.metal tile loop (oracle/pm_oracle.c, pinned to the reference's own "
                                     "shader by tests/test_ref_pin.py), OpenMP over tile rows", "sample": sample,
                   "requested_steps": args.steps, "requested_warmup": args.warmup},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import __graft_entry__ as ge
    pm = ge.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    size = args.size
    nty = (size + 15) // 16

    # ---- scene: encoded once on rank 0, broadcast once over NCCL/NVLink, adopted from device memory ----
    if rank == 0:
        scene_np = scene_for(pm, args)
        n = torch.tensor([scene_np.size], dtype=torch.int64, device="cuda")
    else:
        scene_np, n = None, torch.zeros(1, dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(n, 0)
    scene_dev = torch.empty(int(n.item()), dtype=torch.uint8, device="cuda")
    if rank == 0:
        scene_dev.copy_(torch.from_numpy(scene_np))
    if world > 1:
        dist.broadcast(scene_dev, 0)
    torch.cuda.synchronize()
    scene_bytes = scene_dev.numel()

    r = pm.PietRenderer(device=local_rank)
    r.drawable_size_will_change(size, size)
    # row-strip shard: every rank derives the same cost-balanced bounds from the broadcast scene (no collective)
    scene_host = scene_dev.cpu().numpy()
    bounds = pm.balanced_strip_bounds(pm.row_costs(scene_host, size, size), world) if (world > 1 and not args.equal_strips) \
        else pm.strip_bounds(nty, world)
    if world > 1:
        r.set_strip(bounds[rank], bounds[rank + 1])
    if args.emulate_world and world == 1:
        ew, eg = [int(x) for x in args.emulate_world.split(":")]
        eb = pm.strip_bounds(nty, ew) if args.equal_strips else pm.balanced_strip_bounds(pm.row_costs(scene_host, size, size), ew)
        r.set_strip(eb[eg], eb[eg + 1])
        bounds = [eb[eg], eb[eg + 1]]
    r.set_scene_device(scene_dev.data_ptr(), scene_bytes)
    strip_rows = r.strip_rows
    fb_bytes = strip_rows * size * 4
    # Timing, decided once for ALL ranks: if ANY rank's strip fits the L2, every rank writes a 256 MiB buffer between
    # frames so that each frame streams its framebuffer to HBM, and frames are timed by per-frame event pairs (the
    # flush is outside them).  Otherwise the frames run back to back inside one event pair.  Inside a frame the
    # kernels overlap either way (programmatic dependent launch, the heavy-tile kernel beside the fill/blend kernel).
    t = torch.tensor([fb_bytes], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    flush = int(t.item()) <= L2_BYTES
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if flush else None
    stream = torch.cuda.ExternalStream(r.stream(), device=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def frames(k):
        """k frames; returns (summed per-frame device ms, stats of the last sync)."""
        total, done, st = 0.0, 0, None
        while done < k:
            chunk = min(256, k - done)  # (the renderer keeps the event pairs of the last 512 frames)
            for _ in range(chunk):
                if flush:
                    with torch.cuda.stream(stream):
                        flush_buf.zero_()
                r.draw()
            st = r.sync()
            total += st.ms_total_sum * (chunk / max(1, st.frames))  # (st.frames > chunk only after a pool-growth re-render)
            done += chunk
        return total, st

    # No flush (every strip larger than L2): the K frames are enqueued back to back without events -- consecutive
    # frames are chained by programmatic dependent launch like the kernels inside a frame -- and timed by one event
    # pair around all of them.  Flush: per-frame event pairs (the kernels inside a frame still overlap), summed.
    r.set_frame_events(2 if flush else 0)
    frames(max(3, args.warmup))
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
    ms_sum, st = frames(args.steps)
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if not flush:
        ms_sum = ev0.elapsed_time(ev1)
    # second pass, every kernel group timed on its own (serial): binning / heavy tiles / fill-blend
    r.set_frame_events(1)
    frames(3)
    _, stk = frames(min(args.steps, 256))
    nk = max(1, stk.frames)
    fine_ms, bin_ms, heavy_ms = stk.ms_fine_sum / nk, stk.ms_bin_sum / nk, stk.ms_heavy_sum / nk
    t = torch.tensor([ms_sum, fine_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, fine_ms_max = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = size * size / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the fill/blend kernel (this rank's strip; max duration over ranks) ----
    # algorithmic bytes: the strip's RGBA8 pixels, stored exactly once, minus the tiles k_heavy stores, plus the scene
    peak, peak_src = measured_peak()
    algo_bytes = fb_bytes - int(stk.n_heavy_tiles) * 1024 + scene_bytes
    achieved = algo_bytes / (fine_ms_max * 1e-3) / 1e9

    # ---- e2e: host scene bytes -> H2D -> plan -> frame -> D2H pixels, through pm_renderer_render_host ----
    r.set_frame_events(0)
    host_scene = torch.empty(scene_bytes, dtype=torch.uint8).pin_memory()
    host_scene.copy_(scene_dev.cpu())
    host_out = torch.empty((strip_rows, size, 4), dtype=torch.uint8).pin_memory()
    hs, ho = host_scene.numpy(), host_out.numpy()
    for _ in range(2):
        r.render_host(hs, ho)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        r.render_host(hs, ho)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    plan_ms = r.sync().ms_plan
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = size * size / float(te[0]) / 1e6

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args),
                "scene_bytes": scene_bytes, "tile": "16x16", "parallelism": "row-strips x%d (%s)" % (world, "equal height" if args.equal_strips or world == 1 else "cost-balanced"),
                "strip_tile_rows": [bounds[g + 1] - bounds[g] for g in range(len(bounds) - 1)],
                "timing": ("per-frame CUDA event pairs on the render stream, summed (kernels of a frame overlapped by programmatic dependent launch)"
                           if flush else "one CUDA event pair around %d back-to-back frames (kernels and consecutive frames overlapped by programmatic "
                           "dependent launch)" % args.steps) + "; max over ranks; the choice is the same on every rank",
                "l2": "256 MiB written between frames, outside the event pairs (a strip of this run fits the 126 MB L2)" if flush
                      else "framebuffer strip %.0f MiB > 126 MB L2 on every rank: every frame streams to HBM" % (fb_bytes / 2**20),
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": fb_bytes,
                    "steps": args.e2e_steps, "call": "pm_renderer_render_host (pinned host buffers): upload, validate, plan, frame, read-back",
                    "plan_ms": plan_ms},
            "gpu_launches": int(st.n_launches) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_fine (fill/blend: stores the framebuffer)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "traffic_note": "not measured in this run; the ncu capture of this workload is profiles/r02d_ncu_summary.json (k_fine: 29.6 MB read + 225.1 MB written per launch)",
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": fine_ms_max,
                         "bin_kernel_ms": bin_ms, "heavy_kernel_ms": heavy_ms, "plan_ms": plan_ms,
                         "note": "kernel times from a second pass with every kernel group timed on its own (no overlap)"},
            "frame_stats": {"overflow_records": st.n_overflow_records, "complex_tiles": st.n_complex_tiles, "heavy_tiles": st.n_heavy_tiles, "tiles": st.n_tiles},
        }
        if world == 1 and not args.no_cpu_baseline and not args.emulate_world:  # (the CPU leg is timed at N=1 only)
            threads = host_threads()
            reps, dt = cpu_frames(scene_host, size, threads, min_seconds=10.0)
            out["cpu_baseline"] = {"value": size * size * reps / dt / 1e6, "unit": "Mpixel/s", "cores": threads, "kind": "port",
                                   "sample": "%d whole frames of the same workload, %.1f s wall on %d threads" % (reps, dt, threads)}
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
