"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same encoded scene.

Bar (BASELINE.json north_star): per-tile item lists and solid colours bit-exact; fp32 RGBA within
1e-5 absolute; RGBA8 within 1 LSB."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5


@pytest.fixture(scope="module")
def renderer(pm):
    r = pm.PietRenderer(device=0)
    yield r
    r.close()


def gpu_render(r, scene, w, h, strip=None):
    r.drawable_size_will_change(w, h)
    if strip:
        r.set_strip(*strip)
    r.init_scene(scene)
    r.draw()
    stats = r.sync()
    out = {"rgba8": r.read_rgba8(), "rgba32f": r.read_rgba32f(), "stats": stats}
    out["offsets"], out["items"], out["solid"] = r.read_tile_items()
    # the debug renders must not disturb the product framebuffer
    assert np.array_equal(out["rgba8"], r.read_rgba8())
    return out


def check(gpu, ref, what):
    assert scenes.items_equal(gpu, ref), "%s: per-tile item lists differ" % what
    d8 = np.abs(gpu["rgba8"].astype(np.int16) - ref["rgba8"].astype(np.int16)).max()
    assert d8 <= 1, "%s: RGBA8 differs by %d LSB" % (what, d8)
    a, b = gpu["rgba32f"], ref["rgba32f"]
    assert np.array_equal(np.isnan(a), np.isnan(b)), "%s: NaN pattern differs" % what
    df = np.nanmax(np.abs(a - b)) if a.size else 0.0
    assert df <= F32_TOL, "%s: fp32 RGBA differs by %g" % (what, df)


CASES = [
    ("rect1_int", lambda pm: (pm.build_scene(pm.SCENE_RECT1, 16, 16, rect=(3, 2, 13, 14)), 16, 16)),
    ("rect1_frac", lambda pm: (pm.build_scene(pm.SCENE_RECT1, 16, 16, rect=(3.25, 2.5, 12.75, 13.5)), 16, 16)),
    ("path_test", lambda pm: (pm.build_scene(pm.SCENE_PATH_TEST, 320, 816), 320, 816)),
    ("cardioid", lambda pm: (pm.build_scene(pm.SCENE_CARDIOID, 2048, 1536), 2048, 1536)),
    ("tiger_1024", lambda pm: (pm.build_scene(pm.SCENE_TIGER, 1024, 1024), 1024, 1024)),
    ("tiger_1000x700", lambda pm: (pm.build_scene(pm.SCENE_TIGER, 1000, 1000), 1000, 700)),
    ("rand_bezier_1024", lambda pm: (pm.build_scene(pm.SCENE_RAND_BEZIER, 1024, 1024, count=400), 1024, 1024)),
    ("glyphs_512", lambda pm: (pm.build_scene(pm.SCENE_GLYPHS, 512, 512, count=2000), 512, 512)),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_scene_matches_oracle(pm, oracle, renderer, name, make):
    scene, w, h = make(pm)
    gpu = gpu_render(renderer, scene, w, h)
    ref = oracle.render(scene, w, h, f32=True, items=True)
    check(gpu, ref, name)


def test_fuzz_knife_edges(pm, oracle):
    for seed in range(120):
        scene, w, h, flags = scenes.fuzz_case(pm, seed)
        r = pm.PietRenderer(device=0, flags=flags)
        try:
            gpu = gpu_render(r, scene, w, h)
        finally:
            r.close()
        ref = oracle.render(scene, w, h, flags=flags, f32=True, items=True)
        check(gpu, ref, "fuzz seed %d" % seed)


def test_fill_rules_extension(pm, oracle):
    """PM_FLAG_FILL_RULES: PietFill.flags bit 0 = even-odd rule (TestApp/PietRender.metal:538-540 gives the formula,
    TestApp/SceneEncoder.h:44 reserves the word), multi-subpath fills (holes) through pm_encoder_fill_subpaths, the tiger with
    one compound Fill item per <path>.  GPU against the oracle with the same extension switched on."""
    r = pm.PietRenderer(device=0, flags=pm.FLAG_FILL_RULES)
    try:
        for seed in range(60):
            scene, w, h = scenes.rules_case(pm, seed)
            check(gpu_render(r, scene, w, h), oracle.render(scene, w, h, flags=pm.FLAG_FILL_RULES, f32=True, items=True), "rules seed %d" % seed)
        for opts in (pm.SCENE_OPT_COMPOUND_FILLS, pm.SCENE_OPT_COMPOUND_FILLS | pm.SCENE_OPT_EVEN_ODD):
            scene = pm.build_scene(pm.SCENE_TIGER, 1024, 1024, options=opts)
            check(gpu_render(r, scene, 1024, 1024), oracle.render(scene, 1024, 1024, flags=pm.FLAG_FILL_RULES, f32=True, items=True), "compound tiger %d" % opts)
    finally:
        r.close()
    # without the flag the word is ignored, as upstream: an even-odd item renders like a nonzero one
    r = pm.PietRenderer(device=0)
    try:
        scene, w, h = scenes.rules_case(pm, 3)
        check(gpu_render(r, scene, w, h), oracle.render(scene, w, h, f32=True, items=True), "rules ignored")
    finally:
        r.close()


def test_exact_srgb_flag_is_tighter(pm, oracle):
    scene, w, h = pm.build_scene(pm.SCENE_TIGER, 512, 512), 512, 512
    r = pm.PietRenderer(device=0, flags=pm.FLAG_EXACT_SRGB)
    try:
        gpu = gpu_render(r, scene, w, h)
    finally:
        r.close()
    ref = oracle.render(scene, w, h, f32=True, items=True)
    check(gpu, ref, "tiger_512 exact sRGB")
    assert np.abs(gpu["rgba32f"] - ref["rgba32f"]).max() <= 2e-6  # fixed-point coverage + powf ULPs


def test_row_strips_equal_full_frame(pm, renderer):
    """Multi-GPU shard property: N contiguous row-strips reproduce the 1-GPU frame byte for byte."""
    w = h = 1536
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    full = gpu_render(renderer, scene, w, h)["rgba8"]
    for n in (2, 3, 8):
        b = pm.strip_bounds((h + 15) // 16, n)
        parts = []
        for g in range(n):
            renderer.drawable_size_will_change(w, h)
            renderer.set_strip(b[g], b[g + 1])
            renderer.init_scene(scene)
            renderer.draw()
            parts.append(renderer.read_rgba8())
        assert np.array_equal(np.concatenate(parts, axis=0), full), "%d strips differ from the full frame" % n


def test_render_host_roundtrip(pm, renderer):
    w = h = 640
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    renderer.drawable_size_will_change(w, h)
    img, stats = renderer.render_host(scene)
    renderer.init_scene(scene)
    renderer.draw()
    assert np.array_equal(img, renderer.read_rgba8())
    assert stats.n_launches == 5 and stats.n_complex_tiles > 0  # k_seg, k_row, k_list, k_heavy, k_fine


def test_scene_device_pointer_path(pm, renderer):
    """set_scene_device: the entry point a rank uses after the NCCL broadcast of the scene."""
    import torch
    w = h = 512
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    renderer.drawable_size_will_change(w, h)
    renderer.init_scene(scene)
    renderer.draw()
    want = renderer.read_rgba8()
    t = torch.from_numpy(scene).cuda()
    torch.cuda.synchronize()
    renderer.set_scene_device(t.data_ptr(), t.numel())
    renderer.draw()
    assert np.array_equal(renderer.read_rgba8(), want)


def test_malformed_scene_is_rejected(pm, renderer):
    scene = pm.build_scene(pm.SCENE_PATH_TEST, 320, 816).copy()
    scene[8 + 8 + 16:8 + 8 + 20].view(np.uint32)[0] = 1 << 30  # points_ix far outside the buffer
    renderer.drawable_size_will_change(64, 64)
    with pytest.raises(pm.PietMetalError) as e:
        renderer.init_scene(scene)
    assert e.value.status == pm.PM_ERR_SCENE_MALFORMED
    # a points_ix that is 4 mod 8 would make the 64-bit point loads of binning fault: the device validator refuses it
    scene = pm.build_scene(pm.SCENE_PATH_TEST, 320, 816).copy()
    pix = int(scene[32:36].view(np.uint32)[0])
    bad = np.concatenate([scene, np.zeros(8, np.uint8)])
    bad[32:36].view(np.uint32)[0] = pix + 4
    with pytest.raises(pm.PietMetalError) as e:
        renderer.init_scene(bad)
    assert e.value.status == pm.PM_ERR_SCENE_MALFORMED
    renderer.init_scene(scene)  # ... and the renderer is still usable
    renderer.draw()
    renderer.read_rgba8()


def test_deep_stacks_overflow_chain(pm, oracle, renderer):
    """Hundreds of records per tile: inline slots + overflow chain + records beyond the fill
    kernel's shared-memory index."""
    for layers in (40, 420):
        scene, w, h = scenes.stacked_scene(pm, layers), 96, 64
        gpu = gpu_render(renderer, scene, w, h)
        ref = oracle.render(scene, w, h, f32=True, items=True)
        check(gpu, ref, "stack of %d" % layers)


def test_record_pool_grows_on_demand(pm, oracle):
    scene, w, h = scenes.stacked_scene(pm, 200), 96, 64
    r = pm.PietRenderer(device=0, scratch_bytes=64 * 32)  # room for 64 overflow records: far too few
    try:
        r.drawable_size_will_change(w, h)
        r.init_scene(scene)
        r.draw()
        stats = r.sync()
        assert stats.retries >= 1 and stats.n_overflow_records > 64
        img = r.read_rgba8()
    finally:
        r.close()
    ref = oracle.render(scene, w, h)["rgba8"]
    assert np.abs(img.astype(np.int16) - ref.astype(np.int16)).max() <= 1


def test_full_size_properties_8192(pm, renderer):
    """BASELINE full size (tiger at 8192^2), checked through size-independent properties: the frame
    equals the concatenation of its row strips, an untouched corner is background white, and a band
    of it matches the oracle rendered for that band only."""
    w = h = 8192
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    renderer.drawable_size_will_change(w, h)
    renderer.init_scene(scene)
    renderer.draw()
    stats = renderer.sync()
    full = renderer.read_rgba8()
    assert stats.n_tiles == 512 * 512
    assert (full[:64, :64] == 255).all()
    import zlib
    crc_full = zlib.crc32(full.tobytes())
    b = pm.strip_bounds(512, 4)
    crc = 0
    for g in range(4):
        renderer.drawable_size_will_change(w, h)
        renderer.set_strip(b[g], b[g + 1])
        renderer.init_scene(scene)
        renderer.draw()
        part = renderer.read_rgba8()
        assert np.array_equal(part, full[b[g] * 16:b[g + 1] * 16])
        crc = zlib.crc32(part.tobytes(), crc)
    assert crc == crc_full


def bands(n_rows, n_bands, rows_per_band):
    """n_bands groups of rows_per_band tile rows spread evenly over the frame (first and last row included)."""
    starts = np.linspace(0, n_rows - rows_per_band, n_bands).astype(int)
    return [(int(y), int(y) + rows_per_band) for y in starts]


def test_full_frame_matches_oracle_8192(pm, oracle, renderer):
    """The headline workload, WHOLE frame: the tiger at 8192^2, every one of its 262,144 tiles -- per-tile item lists
    and solid colours bit-exact, fp32 RGBA within 1e-5, RGBA8 within 1 LSB -- in eight strips of 64 tile rows to
    bound the memory of the debug read-backs."""
    w = h = 8192
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    for y0 in range(0, 512, 64):
        gpu = gpu_render(renderer, scene, w, h, strip=(y0, y0 + 64))
        ref = oracle.render(scene, w, h, tile_y0=y0, tile_y1=y0 + 64, f32=True, items=True)
        check(gpu, ref, "tiger 8192 rows %d..%d" % (y0, y0 + 64))


def test_tiger_16384_bands_match_oracle(pm, oracle, renderer):
    """The size the strong-scaling target is quoted on: eight tile rows of the 16384^2 tiger (four bands of two,
    top edge, two through the drawing, bottom edge) against the oracle."""
    w = h = 16384
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    for y0, y1 in bands(1024, 4, 2):
        gpu = gpu_render(renderer, scene, w, h, strip=(y0, y1))
        ref = oracle.render(scene, w, h, tile_y0=y0, tile_y1=y1, f32=True, items=True)
        check(gpu, ref, "tiger 16384 rows %d..%d" % (y0, y1))


def test_config4_bands_match_oracle(pm, oracle, renderer):
    """BASELINE config 4 (10k random filled Bezier paths, 8192^2): 32 tile rows in eight bands spread over the frame."""
    w = h = 8192
    scene = pm.build_scene(pm.SCENE_RAND_BEZIER, w, h)
    for y0, y1 in bands(512, 8, 4):
        gpu = gpu_render(renderer, scene, w, h, strip=(y0, y1))
        ref = oracle.render(scene, w, h, tile_y0=y0, tile_y1=y1, f32=True, items=True)
        check(gpu, ref, "rand_bezier 8192 rows %d..%d" % (y0, y1))


def test_config5_bands_match_oracle(pm, oracle, renderer):
    """BASELINE config 5 (100k glyph-like outlines, 4096^2): 32 tile rows in eight bands spread over the frame."""
    w = h = 4096
    scene = pm.build_scene(pm.SCENE_GLYPHS, w, h)
    for y0, y1 in bands(256, 8, 4):
        gpu = gpu_render(renderer, scene, w, h, strip=(y0, y1))
        ref = oracle.render(scene, w, h, tile_y0=y0, tile_y1=y1, f32=True, items=True)
        check(gpu, ref, "glyphs 4096 rows %d..%d" % (y0, y1))


def test_dense_tiles_256(pm, oracle, renderer):
    """The whole tiger squeezed into 256 x 256 pixels (the smoke frame): thousands of records per tile, every tile
    drawn by k_heavy -- sorted path and overflow block chains."""
    w = h = 256
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    gpu = gpu_render(renderer, scene, w, h)
    ref = oracle.render(scene, w, h, f32=True, items=True)
    check(gpu, ref, "tiger 256")
    assert gpu["stats"].n_heavy_tiles > 100


def test_extreme_density_64(pm, oracle, renderer):
    """The tiger in 64 x 64 pixels: more than 4096 records in a tile, i.e. beyond k_heavy's shared-memory sort (one
    pass over all of the tile's records per item instead)."""
    w = h = 64
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    gpu = gpu_render(renderer, scene, w, h)
    ref = oracle.render(scene, w, h, f32=True, items=True)
    check(gpu, ref, "tiger 64")
    assert gpu["stats"].n_overflow_records > 16 * 4096


def test_empty_and_tiny_surfaces(pm, oracle, renderer):
    """Edge cases: a scene without items (every tile is background white), a 1x1 and a 17x33 surface (ragged
    right and bottom tiles)."""
    empty = np.zeros(8, np.uint8)
    empty[4:8].view(np.uint32)[0] = 8  # n_items = 0, items_ix = 8
    for w, h in ((16, 16), (1, 1), (17, 33)):
        gpu = gpu_render(renderer, empty, w, h)
        assert gpu["rgba8"].shape == (h, w, 4) and (gpu["rgba8"] == 255).all()
        assert len(gpu["items"]) == 0 and (gpu["solid"] == 0xffffffff).all()
    for w, h in ((1, 1), (17, 33), (33, 17)):
        scene = pm.build_scene(pm.SCENE_RECT1, w, h, rect=(0.5, 0.25, w - 0.25, h - 0.5))
        gpu = gpu_render(renderer, scene, w, h)
        ref = oracle.render(scene, w, h, f32=True, items=True)
        check(gpu, ref, "rect on %dx%d" % (w, h))


def test_frames_without_events_are_identical(pm, renderer):
    """pm_renderer_set_frame_events(0): the frame's kernels are chained by programmatic dependent launch and
    consecutive frames overlap their launches; the pixels must not change, frame after frame."""
    w = h = 1024
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    renderer.drawable_size_will_change(w, h)
    renderer.init_scene(scene)
    renderer.draw()
    want = renderer.read_rgba8()
    renderer.set_frame_events(False)
    try:
        for _ in range(25):
            renderer.draw()
        st = renderer.sync()
        assert st.frames == 0 and st.n_complex_tiles > 0
        assert np.array_equal(renderer.read_rgba8(), want)
        # a debug read-back in between (it renders a frame of its own) and more overlapped frames
        f32 = renderer.read_rgba32f()
        assert f32.shape == (h, w, 4)
        for _ in range(3):
            renderer.draw()
        assert np.array_equal(renderer.read_rgba8(), want)
    finally:
        renderer.set_frame_events(True)
    renderer.draw()
    st = renderer.sync()
    assert st.frames == 1 and st.ms_total > 0
    assert np.array_equal(renderer.read_rgba8(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("n_frames", [2, 7])
def test_overlapped_frames_large(pm, renderer, n_frames):
    """Back-to-back frames without events (kernels and frames chained by programmatic dependent launch) at a size
    where the kernels' tails are long: after an even and an odd number of frames, pixels and per-tile item lists are
    those of a single synchronous frame."""
    w = h = 4096
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    renderer.drawable_size_will_change(w, h)
    renderer.init_scene(scene)
    renderer.draw()
    want = renderer.read_rgba8()
    want_off, want_items, want_solid = renderer.read_tile_items()
    renderer.set_frame_events(False)
    try:
        for _ in range(n_frames):
            renderer.draw()
        renderer.sync()
        assert np.array_equal(renderer.read_rgba8(), want)
        off, items, solid = renderer.read_tile_items()
        assert np.array_equal(off, want_off) and np.array_equal(items, want_items) and np.array_equal(solid, want_solid)
    finally:
        renderer.set_frame_events(True)


def test_balanced_strips_equal_full_frame(pm, renderer):
    """Cost-balanced (unequal) row strips of the tiger and of the glyph scene reproduce the full frame byte for byte."""
    for kind, size, count in ((pm.SCENE_TIGER, 2048, 0), (pm.SCENE_GLYPHS, 1024, 6000)):
        scene = pm.build_scene(kind, size, size, count=count)
        full = gpu_render(renderer, scene, size, size)["rgba8"]
        b = pm.balanced_strip_bounds(pm.row_costs(scene, size, size), 8)
        assert len(set(b[i + 1] - b[i] for i in range(8))) > 1  # really unequal
        parts = []
        for g in range(8):
            renderer.drawable_size_will_change(size, size)
            renderer.set_strip(b[g], b[g + 1])
            renderer.init_scene(scene)
            renderer.draw()
            parts.append(renderer.read_rgba8())
        assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_many_records_per_tile_extension_and_chain(pm, oracle, renderer):
    """Tiles with 17..63 records (extension block) and beyond 64 (chain): a fan of thin polygons through one
    point, drawn with translucent colours so that nothing is rewound away."""
    w, h = 64, 48
    for n_blades in (20, 70, 150):
        enc = pm.Encoder(1 << 20)
        enc.begin_group(n_blades)
        for k in range(n_blades):
            a = np.pi * k / n_blades
            dx, dy = np.cos(a), np.sin(a)
            cx, cy = 24.3, 20.7
            pts = [(cx - 40 * dx - 0.6 * dy, cy - 40 * dy + 0.6 * dx), (cx + 40 * dx - 0.6 * dy, cy + 40 * dy + 0.6 * dx),
                   (cx + 40 * dx + 0.6 * dy, cy + 40 * dy - 0.6 * dx), (cx - 40 * dx + 0.6 * dy, cy - 40 * dy - 0.6 * dx)]
            enc.fill(np.array(pts), 0x20406080 + (k << 8))
        enc.end_group()
        scene = enc.bytes()
        gpu = gpu_render(renderer, scene, w, h)
        ref = oracle.render(scene, w, h, f32=True, items=True)
        check(gpu, ref, "fan of %d" % n_blades)
