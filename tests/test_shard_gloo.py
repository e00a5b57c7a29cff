"""Multi-GPU host logic on CPU: world_size 2 over gloo.

The N>1 path is: rank 0 encodes the scene, broadcasts its length and bytes once, every rank renders
its contiguous row-strip with no further collective, rank 0 gathers the strips (off the hot path).
Here the per-rank renderer is the host harness of the product's logic headers (no GPU in this test);
the gathered frame must equal the single-process frame byte for byte."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, size):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as ge
    import oracle_api
    pm = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = torch.zeros(1, dtype=torch.int64)
    if rank == 0:
        scene = torch.from_numpy(pm.build_scene(pm.SCENE_TIGER, size, size))
        n[0] = scene.numel()
    dist.broadcast(n, 0)
    if rank != 0:
        scene = torch.empty(int(n[0]), dtype=torch.uint8)
    dist.broadcast(scene, 0)                       # the one collective: scene bytes, once
    # every rank derives the same cost-balanced strips from the broadcast scene
    bounds = pm.balanced_strip_bounds(pm.row_costs(scene.numpy(), size, size), world)
    strip = oracle_api.harness_render(scene.numpy(), size, size, tile_y0=bounds[rank], tile_y1=bounds[rank + 1])["rgba8"]
    rows = [(bounds[g + 1] - bounds[g]) * 16 for g in range(world)]
    if rank == 0:
        parts = [torch.empty((rows[g], size, 4), dtype=torch.uint8) for g in range(world)]
        parts[0] = torch.from_numpy(strip)
        for g in range(1, world):
            dist.recv(parts[g], src=g)
        np.save(os.path.join(out_dir, "frame.npy"), torch.cat(parts, 0).numpy())
    else:
        dist.send(torch.from_numpy(strip), dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_strips_reproduce_the_frame(pm, oracle, tmp_path):
    size, world, port = 272, 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), size), nprocs=world, join=True)
    frame = np.load(os.path.join(str(tmp_path), "frame.npy"))
    scene = pm.build_scene(pm.SCENE_TIGER, size, size)
    assert np.array_equal(frame, oracle.harness_render(scene, size, size)["rgba8"])
    assert np.abs(frame.astype(np.int16) - oracle.render(scene, size, size)["rgba8"].astype(np.int16)).max() <= 1


def test_strip_bounds_partition(pm):
    for rows in (1, 7, 64, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            b = pm.strip_bounds(rows, world)
            assert b[0] == 0 and b[-1] == rows and all(b[i] <= b[i + 1] for i in range(world))
            assert max(b[i + 1] - b[i] for i in range(world)) - min(b[i + 1] - b[i] for i in range(world)) <= 1


def test_balanced_strips_properties(pm):
    """pm_balance_strips: contiguous, non-empty, covering; never worse than equal-height strips; optimal on
    small cases (brute force)."""
    import itertools
    rng = np.random.default_rng(7)
    for rows, world in ((1, 1), (5, 5), (9, 3), (12, 4), (64, 8), (512, 8)):
        for trial in range(4):
            cost = rng.random(rows).astype(np.float32) * (10.0 if trial % 2 else 1.0) + (0.0 if trial < 2 else 5.0)
            b = pm.balanced_strip_bounds(cost, world)
            assert b[0] == 0 and b[-1] == rows and all(b[i] < b[i + 1] for i in range(world))
            worst = max(cost[b[i]:b[i + 1]].sum() for i in range(world))
            eq = pm.strip_bounds(rows, world)
            if all(eq[i] < eq[i + 1] for i in range(world)):
                assert worst <= max(cost[eq[i]:eq[i + 1]].sum() for i in range(world)) * (1 + 1e-5)
            if rows <= 12:
                best = min(max(cost[c0:c1].sum() for c0, c1 in zip((0,) + cuts, cuts + (rows,)))
                           for cuts in itertools.combinations(range(1, rows), world - 1))
                assert worst <= best * (1 + 1e-5)
    with pytest.raises(pm.PietMetalError):
        pm.balanced_strip_bounds(np.ones(3, np.float32), 4)


def test_row_costs_follow_the_drawing(pm):
    """pm_scene_row_costs: every row costs at least the per-tile term; the tiger's middle rows cost several times
    the cheapest row; the balanced 8-way split of the 8192^2 tiger is within 2 % of even."""
    size = 2048
    scene = pm.build_scene(pm.SCENE_TIGER, size, size)
    cost = pm.row_costs(scene, size, size)
    assert cost.shape == (size // 16,) and (cost > 0).all()
    assert cost.min() >= 0.25 * (size // 16) and cost[len(cost) // 2] > 3 * cost.min()
    b = pm.balanced_strip_bounds(cost, 8)
    parts = np.array([cost[b[i]:b[i + 1]].sum() for i in range(8)])
    assert parts.max() / parts.mean() < 1.06
    rect = pm.build_scene(pm.SCENE_RECT1, 64, 64, rect=(3, 2, 13, 14))
    c = pm.row_costs(rect, 64, 64)
    assert c[0] > c[1] == c[2] == c[3]
