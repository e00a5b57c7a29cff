"""The product's binning / fill logic (pm_tile_logic.h, pm_pixel_logic.h), replayed on the CPU by
tests/native/pm_host_harness.cpp, against the oracle's literal per-tile loop.

Per-tile item lists, backdrops and solid colours must be bit-exact (the restructuring into per
(segment, tile row) work with bisected backdrop suffixes is claimed to be exact); pixels differ only
by the 2^-24 fixed-point rounding of the coverage sums."""
import numpy as np
import pytest

import scenes

F32_TOL = 2e-6


def compare(oracle, scene, w, h, flags=0):
    o = oracle.render(scene, w, h, flags=flags, f32=True, items=True)
    g = oracle.harness_render(scene, w, h, flags=flags, f32=True, items=True)
    assert scenes.items_equal(o, g), "per-tile item lists differ"
    a, b = o["rgba32f"], g["rgba32f"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) <= F32_TOL
    assert np.abs(o["rgba8"].astype(np.int16) - g["rgba8"].astype(np.int16)).max() <= 1


@pytest.mark.parametrize("name", ["rect", "path_test", "cardioid", "tiger", "tiger_odd", "bezier", "glyphs", "stack"])
def test_scene(pm, oracle, name):
    if name == "rect":
        compare(oracle, pm.build_scene(pm.SCENE_RECT1, 64, 48, rect=(3.25, 2.5, 52.75, 43.5)), 64, 48)
    elif name == "path_test":
        compare(oracle, pm.build_scene(pm.SCENE_PATH_TEST, 320, 816), 320, 816)
    elif name == "cardioid":
        compare(oracle, pm.build_scene(pm.SCENE_CARDIOID, 1024, 768, scale=0.5), 1024, 768)
    elif name == "tiger":
        compare(oracle, pm.build_scene(pm.SCENE_TIGER, 768, 768), 768, 768)
    elif name == "tiger_odd":
        compare(oracle, pm.build_scene(pm.SCENE_TIGER, 1000, 1000), 1000, 700, flags=1)
    elif name == "bezier":
        compare(oracle, pm.build_scene(pm.SCENE_RAND_BEZIER, 1024, 1024, count=300), 1024, 1024)
    elif name == "glyphs":
        compare(oracle, pm.build_scene(pm.SCENE_GLYPHS, 512, 512, count=1500), 512, 512)
    else:
        compare(oracle, scenes.stacked_scene(pm, 120), 96, 64)


def test_fuzz_knife_edges(pm, oracle):
    """Vertices on tile corners / edges, horizontal and vertical segments, long diagonals across
    several 256-px strips: every cull decision that can sit on a knife edge."""
    for seed in range(250):
        scene, w, h, flags = scenes.fuzz_case(pm, seed)
        compare(oracle, scene, w, h, flags)


def test_fuzz_wide_frames(pm, oracle):
    for seed in range(40):
        rng = np.random.default_rng(9000 + seed)
        w, h = int(rng.choice([1100, 2100])), int(rng.choice([48, 100]))
        scene = scenes.random_scene(pm, 500 + seed, w, h, int(rng.integers(3, 14)), scenes.FUZZ_MODES[seed % 4])
        compare(oracle, scene, w, h, flags=seed % 2)


def test_row_strips(pm, oracle):
    scene = pm.build_scene(pm.SCENE_TIGER, 512, 512)
    full = oracle.harness_render(scene, 512, 512)["rgba8"]
    parts = [oracle.harness_render(scene, 512, 512, tile_y0=a, tile_y1=b)["rgba8"] for a, b in ((0, 11), (11, 12), (12, 32))]
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_fill_rules_extension(pm, oracle):
    """PM_FLAG_FILL_RULES (extension): even-odd items and two-subpath fills through the device's binning / fill logic
    (replayed on the CPU) against the oracle's literal loop with the same extension switched on."""
    for seed in range(60):
        scene, w, h = scenes.rules_case(pm, seed)
        compare(oracle, scene, w, h, flags=pm.FLAG_FILL_RULES)
    for opts in (pm.SCENE_OPT_COMPOUND_FILLS, pm.SCENE_OPT_COMPOUND_FILLS | pm.SCENE_OPT_EVEN_ODD):
        scene = pm.build_scene(pm.SCENE_TIGER, 512, 512, options=opts)
        compare(oracle, scene, 512, 512, flags=pm.FLAG_FILL_RULES)
