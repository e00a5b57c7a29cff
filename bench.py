#!/usr/bin/env python3
"""bench.py -- Mpixel/s of the compute render path on Ghostscript_Tiger at 8192x8192 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one frame: the encoded scene (resident in HBM) -> binning kernels (k_seg, k_row) -> fill/blend
kernel (k_fine) -> RGBA8 framebuffer in HBM.  With N GPUs the frame's tile rows are sharded into N contiguous
row-strips (strong scaling: the frame is fixed); the scene is broadcast once over NCCL before the
timed region and there is no collective per frame.  Rank 0 prints ONE JSON line.

  value      whole-frame Mpixel/s, K frames back to back (launches overlapped, no per-frame events),
             device-timed with CUDA events around the K frames, max over ranks
  e2e        the same metric through the C-ABI call pm_renderer_render_host: scene bytes in pinned
             host memory -> H2D -> frame -> D2H of the strip's pixels into pinned host memory
  roofline   fill/blend kernel: algorithmic bytes (4*W*H_strip + scene) / its CUDA-event duration (a second
             pass of K frames with per-frame events on the render stream), against the measured HBM copy
             bandwidth in MEASURED_PEAKS.json
  cpu_baseline   the oracle (scalar port of the reference's tile loop) on the host cores, on a
             bounded band of tile rows of the same frame

--impl reference times that CPU port alone (the reference itself is Metal/Rust/Objective-C and
cannot be built here), rank 0 only.
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
    ap.add_argument("--cpu-rows", type=int, default=0, help="tile rows in the CPU sample (0 = auto, ~10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frame-events", action="store_true", help="keep per-frame CUDA events in the headline pass (no launch overlap)")
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


def cpu_sample(scene, size, rows_hint, min_seconds=8.0):
    """Time the oracle (all host cores) on the same frame: the whole frame, or a centred band of
    `rows_hint` tile rows, repeated until about `min_seconds` of wall time have gone by."""
    import oracle_api
    threads = host_threads()
    nty = (size + 15) // 16
    rows = rows_hint if rows_hint > 0 else nty
    y0 = max(0, nty // 2 - rows // 2)
    y1 = min(nty, y0 + rows)
    px = (min(y1 * 16, size) - y0 * 16) * size
    oracle_api.render(scene, size, size, tile_y0=y0, tile_y1=min(y1, y0 + 2), threads=threads)  # warm the thread pool
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle_api.render(scene, size, size, tile_y0=y0, tile_y1=y1, threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= 200:
            break
    desc = "tile rows %d..%d of %d x %d repetitions, %.1f s wall on %d threads" % (y0, y1, nty, reps, dt, threads)
    return px * reps / dt / 1e6, desc, threads


def run_reference(args, rank):
    """--impl reference: the CPU port of the reference's tile loop on this box's host cores (the
    reference itself -- Metal kernels, Rust feed -- cannot be built here)."""
    if rank != 0:
        return
    import __graft_entry__ as ge
    import oracle_api
    pm = ge.load_package()
    scene = scene_for(pm, args)
    threads = host_threads()
    nty = (args.size + 15) // 16
    # each step is a bounded sample (a centred band of tile rows) so that steps + warmup end within minutes
    t = time.perf_counter()
    oracle_api.render(scene, args.size, args.size, tile_y0=nty // 2, tile_y1=nty // 2 + 2, threads=threads)
    per_row = (time.perf_counter() - t) / 2
    budget = 120.0 / max(1, args.steps + args.warmup)
    rows = int(max(1, min(nty, budget / max(per_row, 1e-6))))
    y0 = max(0, nty // 2 - rows // 2)
    y1 = min(nty, y0 + rows)
    for _ in range(args.warmup):
        oracle_api.render(scene, args.size, args.size, tile_y0=y0, tile_y1=y1, threads=threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        oracle_api.render(scene, args.size, args.size, tile_y0=y0, tile_y1=y1, threads=threads)
    dt = time.perf_counter() - t
    px = (min(y1 * 16, args.size) - y0 * 16) * args.size
    value = px * args.steps / dt / 1e6
    sample = "tile rows %d..%d of %d per step, %d threads" % (y0, y1, nty, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Ghostscript_Tiger %dx%d" % (args.size, args.size) if args.scene == "tiger" else "%s %dx%d" % (args.scene, args.size, args.size),
                   "implementation": "CPU port of the PietRender.metal tile loop (oracle/pm_oracle.c), OpenMP", "sample": sample},
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
    flush = fb_bytes <= L2_BYTES  # the strip would stay L2-resident between frames: flush L2 between timed frames
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if flush else None
    stream = torch.cuda.ExternalStream(r.stream(), device=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def frames(k):
        """k frames; returns (device ms summed over the frames, fill-kernel ms summed)."""
        if not flush:
            for _ in range(k):
                r.draw()
            st = r.sync()
            scale = k / max(1, st.frames)  # per-frame event pairs kept for the last <= 512 frames (+ any pool-growth re-render)
            return st.ms_total_sum * scale, st.ms_fine_sum * scale, st
        total = fine = 0.0
        st = None
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
            r.draw()
            st = r.sync()
            total += st.ms_total
            fine += st.ms_fine
        return total, fine, st

    # Headline pass: no per-frame events, so that the frame's kernels (and consecutive frames) overlap their
    # launches; the fill kernel's own duration is measured in a second pass with per-frame events.
    overlap = (not flush) and not args.frame_events
    r.set_frame_events(not overlap)
    frames(max(3, args.warmup))
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
    ms_sum, ms_fine_sum, st = frames(args.steps)
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    wall_ms = ev0.elapsed_time(ev1)
    # back-to-back frames: the event bracket is the step time; with L2 flushes in between, the sum of the
    # per-frame event pairs is (the flush is not part of a step)
    if overlap:  # second pass, per-frame CUDA events on the render stream: binning / fill kernel times
        r.set_frame_events(True)
        frames(3)
        ms_sum, ms_fine_sum, st = frames(args.steps)
    my_ms = ms_sum if flush else wall_ms
    t = torch.tensor([my_ms, ms_fine_sum], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_fine_total = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = size * size / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the fill/blend kernel (this rank's strip; max duration over ranks) ----
    peak, peak_src = measured_peak()
    fine_ms = ms_fine_total / args.steps
    algo_bytes = fb_bytes + scene_bytes
    achieved = algo_bytes / (fine_ms * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch of the fill kernel from the committed ncu capture of this very workload
    if world == 1 and size == 8192 and args.scene == "tiger":
        try:
            import glob
            with open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_fine_traffic.json")))[-1]) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: host scene bytes -> H2D -> frame -> D2H pixels, through pm_renderer_render_host ----
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
                "workload": "Ghostscript_Tiger %dx%d" % (size, size) if args.scene == "tiger" else "%s %dx%d" % (args.scene, size, size),
                "scene_bytes": scene_bytes, "tile": "16x16", "parallelism": "row-strips x%d (%s)" % (world, "equal height" if args.equal_strips or world == 1 else "cost-balanced"),
                "strip_tile_rows": [bounds[g + 1] - bounds[g] for g in range(world)],
                "l2": ("flushed between timed frames (strip %.0f MiB <= L2)" % (fb_bytes / 2**20)) if flush
                      else "framebuffer strip %.0f MiB > 126 MB L2: every frame streams to HBM" % (fb_bytes / 2**20),
                "timing": "sum of per-frame CUDA event pairs" if flush else "CUDA events around %d back-to-back frames" % args.steps,
                "launch_overlap": "programmatic dependent launch between the frame's kernels and between frames" if overlap else "none (per-frame events)",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": fb_bytes,
                    "steps": args.e2e_steps, "call": "pm_renderer_render_host (pinned host buffers)"},
            "gpu_launches": int(st.n_launches) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_fine (fill/blend)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": fine_ms,
                         "bin_kernel_ms": st.ms_bin_sum / max(1, st.frames), "heavy_kernel_ms": st.ms_heavy_sum / max(1, st.frames)},
            "frame_stats": {"overflow_records": st.n_overflow_records, "complex_tiles": st.n_complex_tiles, "heavy_tiles": st.n_heavy_tiles, "tiles": st.n_tiles},
        }
        if world == 1 and not args.no_cpu_baseline:  # (the CPU leg is timed at N=1 only)
            v, desc, threads = cpu_sample(scene_host, size, args.cpu_rows)
            out["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": desc}
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
