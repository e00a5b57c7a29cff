"""Known-answer tests that pin the oracle (the reference ships none: SURVEY.md section 4).

Every expected value here is derived by hand from TestApp/PietRender.metal, not from running code."""
import numpy as np

import scenes

CMD_END, CMD_FILL, CMD_FILLEDGE, CMD_DRAWFILL, CMD_SOLID, CMD_BAIL = 1, 4, 6, 7, 8, 9


def srgb_encode(v):
    return 12.92 * v if v < 0.0031308 else 1.055 * v ** (1 / 2.4) - 0.055


def srgb_decode(b):
    c = b / 255.0
    return c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4


def test_config1_integer_rect_coverage_is_exactly_one_inside(pm, oracle):
    """BASELINE config 1: one 16x16 tile, rect (3,2)-(13,14), colour 0x3366ccff."""
    scene = pm.build_scene(pm.SCENE_RECT1, 16, 16, rect=(3, 2, 13, 14))
    out = oracle.render(scene, 16, 16, f32=True, items=True)
    assert out["offsets"].tolist() == [0, 1]
    assert out["items"].tolist() == [(0, 0, 0)]          # item 0 draws, backdrop 0 at the tile's corner
    assert out["solid"].tolist() == [0]
    img = out["rgba8"]
    assert (img[2:14, 3:13] == [0x33, 0x66, 0xCC, 0xFF]).all()
    outside = np.ones((16, 16), bool); outside[2:14, 3:13] = False
    assert (img[outside] == 255).all()
    f = out["rgba32f"]
    want = [srgb_encode(srgb_decode(b)) for b in (0x33, 0x66, 0xCC)]
    assert np.abs(f[5, 5, :3] - want).max() < 1e-6 and f[5, 5, 3] == 1.0


def test_config1_fractional_rect_edges_are_exact_fractions(pm, oracle):
    """Fractional rect (3.25,2.5)-(12.75,13.5): edge pixels get area fractions 0.75 / 0.5 exactly
    (the pixel sample is its integer corner, coverage is the exact trapezoid area, metal:508-528)."""
    scene = pm.build_scene(pm.SCENE_RECT1, 16, 16, rect=(3.25, 2.5, 12.75, 13.5), rgba=0x000000FF)
    f = oracle.render(scene, 16, 16, f32=True)["rgba32f"][:, :, 0]
    def alpha_of(v):  # black over white: linear value = 1 - alpha; invert the sRGB encode
        lin = v / 12.92 if v < 12.92 * 0.0031308 else ((v + 0.055) / 1.055) ** 2.4
        return 1.0 - lin
    assert abs(alpha_of(f[5, 5]) - 1.0) < 1e-6      # interior
    assert abs(alpha_of(f[2, 5]) - 0.5) < 1e-6      # top edge: y in [2.5, 3) -- window fractions are exact
    assert abs(alpha_of(f[13, 5]) - 0.5) < 1e-6     # bottom edge
    # An exactly vertical edge goes through the reference's 1e-6 fudge (metal:519-524): the area is
    # (b + (d^2 - c^2)/2 - xmin) / 1e-6 with catastrophic cancellation in fp32, so 0.75 comes out as
    # 0.75 +- 0.05.  That noise is the reference's; the oracle reproduces the formula literally.
    assert abs(alpha_of(f[5, 3]) - 0.75) < 0.06     # left edge: x in [3.25, 4) covered
    assert abs(alpha_of(f[5, 12]) - 0.75) < 0.06    # right edge: x in [12, 12.75)
    assert abs(alpha_of(f[2, 3]) - 0.375) < 0.03    # corner: 0.75 * 0.5
    assert abs(f[0, 0] - 1.0) < 1e-6 and abs(f[5, 14] - 1.0) < 1e-6   # 1.055 * 1^(1/2.4) - 0.055 in fp32


def test_command_stream_of_a_tile_right_of_the_shape(pm, oracle):
    """A tile wholly inside a rect that starts left of it: no Fill commands, backdrop != 0 -> opaque
    Solid rewinds the list and the tile Bails (metal:127-151)."""
    scene = pm.build_scene(pm.SCENE_RECT1, 64, 48, rect=(3.25, 2.5, 60.0, 43.5))
    cmds, solid = oracle.tile_cmds(scene, 2, 1)
    assert solid == 0xFFCC6633
    assert cmds["tag"].tolist() == [CMD_BAIL]
    # the left-most tile crosses the left edge: Fill commands, DrawFill with backdrop 0, End
    cmds, solid = oracle.tile_cmds(scene, 0, 1)
    assert solid == 0
    tags = cmds["tag"].tolist()
    assert tags[-2:] == [CMD_DRAWFILL, CMD_END] and CMD_FILL in tags
    assert cmds["body"][tags.index(CMD_DRAWFILL)][0] == 0
    # a tile to the right of x = 16 sees the edge segments crossing its left edge: FillEdge + Fill
    cmds, _ = oracle.tile_cmds(scene, 1, 0)
    assert CMD_FILLEDGE in cmds["tag"].tolist()


def test_winding_sign_symmetry(pm, oracle):
    """Nonzero rule: reversing a polygon's direction leaves every pixel unchanged (metal:537)."""
    pts = np.array([(5.3, 4.1), (58.2, 9.7), (49.9, 40.2), (20.5, 44.4), (9.1, 30.0)])
    imgs = []
    for p in (pts, pts[::-1]):
        enc = pm.Encoder(4096); enc.begin_group(1); enc.fill(p, 0x204080FF); enc.end_group()
        imgs.append(oracle.render(enc.bytes(), 64, 48, f32=True)["rgba32f"])
    assert np.abs(imgs[0] - imgs[1]).max() < 1e-6


def test_coverage_sums_to_polygon_area(pm, oracle):
    """Sum of coverage over the pixels == polygon area (exact-area rasterisation, metal:508-528)."""
    pts = np.array([(5.3, 4.1), (58.2, 9.7), (49.9, 40.2), (20.5, 44.4), (9.1, 30.0)])
    area = 0.5 * abs(np.dot(pts[:, 0], np.roll(pts[:, 1], -1)) - np.dot(pts[:, 1], np.roll(pts[:, 0], -1)))
    enc = pm.Encoder(4096); enc.begin_group(1); enc.fill(pts, 0x000000FF); enc.end_group()
    f = oracle.render(enc.bytes(), 64, 48, f32=True)["rgba32f"][:, :, 0]
    # black over white: linear value = 1 - coverage; invert the sRGB encode
    lin = np.where(f < 12.92 * 0.0031308, f / 12.92, ((f + 0.055) / 1.055) ** 2.4)
    assert abs((1.0 - lin).sum() - area) < 1e-2


def test_tile_translation_invariance(pm, oracle):
    """Shifting the scene by one tiler group (256 px, 32 px) shifts the image; the pixels agree up
    to the fp32 rounding of the shifted coordinates."""
    rng = np.random.default_rng(5)
    pts = rng.uniform(10, 200, (7, 2))
    imgs = []
    for dx, dy in ((0, 0), (256, 32)):
        enc = pm.Encoder(4096); enc.begin_group(2)
        enc.fill(pts + (dx, dy), 0x8040C0FF)
        enc.polyline(pts[:4] + (dx, dy), 0x102030C0, 3.0)
        enc.end_group()
        imgs.append(oracle.render(enc.bytes(), 512, 272)["rgba8"])
    d = np.abs(imgs[0][:240, :256].astype(np.int16) - imgs[1][32:272, 256:512].astype(np.int16))
    assert d.max() <= 1 and (d != 0).mean() < 0.01


def test_opaque_cover_rewinds_and_translucent_cover_is_dropped_on_bail(pm, oracle):
    """Quirks 1 and 2 of SURVEY.md 8(a): an opaque full cover discards earlier items; a translucent
    full cover over a solid tile is lost because the tile still Bails."""
    big = np.array([(-50.0, -50.0), (300.0, -50.0), (300.0, 300.0), (-50.0, 300.0)])
    enc = pm.Encoder(8192); enc.begin_group(3)
    enc.fill(np.array([(20.0, 20.0), (40.0, 22.0), (30.0, 44.0)]), 0xFF0000FF)   # hidden triangle
    enc.fill(big, 0x00FF00FF)                                                    # opaque cover
    enc.fill(big, 0x0000FF80)                                                    # translucent cover
    enc.end_group()
    out = oracle.render(enc.bytes(), 64, 64, items=True)
    t = 1 * 4 + 1  # tile (1,1) holds the triangle
    lst = out["items"][out["offsets"][t]:out["offsets"][t + 1]].tolist()
    assert lst == [(1, 0, 1), (2, 0, 1)]                  # list restarts at the opaque cover
    assert out["solid"][t] == 0xFF00FF00                  # ... and the tile Bails with its colour
    assert (out["rgba8"] == [0, 255, 0, 255]).all()       # the translucent blue layer is dropped


def test_empty_scene_is_white(pm, oracle):
    enc = pm.Encoder(64); enc.begin_group(0); enc.end_group()
    out = oracle.render(enc.bytes(), 40, 24, items=True)
    assert (out["rgba8"] == 255).all() and (out["solid"] == 0xFFFFFFFF).all() and len(out["items"]) == 0


def test_poly_precull_quirk_and_fix_flag_differ_only_for_strokes(pm, oracle):
    """Quirk 10: the group-level polyline pre-cull uses the voting lane's tile row; the fix flag uses
    the consuming tile's row.  Fills are unaffected."""
    scene = pm.build_scene(pm.SCENE_TIGER, 256, 256)
    a = oracle.render(scene, 256, 256, items=True)
    b = oracle.render(scene, 256, 256, flags=1, items=True)
    kinds = scene[int(scene[4:8].view(np.uint32)[0]):].view(np.uint32)[::8]
    def fills(o):
        return [tuple(x) for x in o["items"].tolist() if kinds[x[0]] == 3]
    assert fills(a) == fills(b)


def test_golden_tiger_128(pm, oracle):
    """Committed fixture (tools/make_golden.py): the tiger at 128x128 through the oracle."""
    import os
    root = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(root, "golden", "tiger_128_rgba8.npy"))
    scene = pm.build_scene(pm.SCENE_TIGER, 128, 128)
    assert np.array_equal(oracle.render(scene, 128, 128)["rgba8"], gold)


# ---------------------------------------------------------------------------------------------------
# Extension: even-odd rule and multi-subpath fills through PietFill.flags
# ---------------------------------------------------------------------------------------------------
PMO_FLAG_FILL_RULES = 4


def _star(pm, flags, rgba=0x000000ff):
    """A pentagram on 64 x 64: its centre pentagon has winding number 2.  (Centre off the tile grid: a vertex exactly on
    a tile's left edge is a case the reference's left-edge split does not handle, with either rule.)"""
    import math
    cx, cy, r = 33.3, 32.7, 28.0
    pts = [(cx + r * math.sin(2 * math.pi * (2 * k) / 5), cy - r * math.cos(2 * math.pi * (2 * k) / 5)) for k in range(5)]
    enc = pm.Encoder(1 << 12)
    enc.begin_group(1)
    enc.fill(np.array(pts), rgba, flags=flags)
    enc.end_group()
    return enc.bytes()


def test_even_odd_rule_empties_the_doubly_wound_centre(pm, oracle):
    eo = _star(pm, pm.FILL_EVEN_ODD)
    nonzero = oracle.render(eo, 64, 64)["rgba8"]                                 # the flag word is ignored by default, as upstream
    assert np.array_equal(nonzero, oracle.render(_star(pm, 0), 64, 64)["rgba8"])
    evenodd = oracle.render(eo, 64, 64, flags=PMO_FLAG_FILL_RULES)["rgba8"]
    assert (nonzero[33, 33, :3] == 0).all() and (evenodd[33, 33, :3] == 255).all()  # centre: winding 2 -> filled / empty
    assert (nonzero[12, 33, :3] == 0).all() and (evenodd[12, 33, :3] == 0).all()    # the top point of the star: winding 1
    assert (evenodd[2, 2, :3] == 255).all()
    # with the extension on, an item without the bit is still filled by the nonzero rule
    assert np.array_equal(oracle.render(_star(pm, 0), 64, 64, flags=PMO_FLAG_FILL_RULES)["rgba8"], nonzero)


def test_subpaths_cut_holes(pm, oracle):
    """One Fill item of two subpaths (pm_encoder_fill_subpaths): an outer square and an inner square wound the other
    way.  Nonzero rule, default renderer: the hole is cut out -- the reference's per-subpath fills would paint it over
    (src/lib.rs:194 'need to deal with subpaths')."""
    # (slightly tilted quadrilaterals: an exactly horizontal edge that crosses a tile boundary gets FillEdge sign 0 in
    # the reference's left-edge split, metal:336-338, and is lost -- with one subpath as with two)
    outer = [(8.3, 8.1), (88.2, 9.4), (87.1, 56.3), (9.2, 55.2)]
    inner = [(30.4, 20.2), (29.6, 44.1), (70.3, 43.2), (69.5, 19.3)]
    enc = pm.Encoder(1 << 12)
    enc.begin_group(1)
    enc.fill_subpaths([outer, inner], 0x203040ff)
    enc.end_group()
    img = oracle.render(enc.bytes(), 96, 64)["rgba8"]
    assert (img[32, 50, :3] == 255).all()          # inside the hole
    assert (img[32, 20, :3] != 255).any()          # in the ring
    assert (img[12, 50, :3] != 255).any()
    assert (img[2, 2, :3] == 255).all()
    # the bridges between the subpaths leave no trace: the picture equals "outer filled, then inner painted white"
    enc2 = pm.Encoder(1 << 12)
    enc2.begin_group(2)
    enc2.fill(np.array(outer), 0x203040ff)
    enc2.fill(np.array(inner), 0xffffffff)
    enc2.end_group()
    two = oracle.render(enc2.bytes(), 96, 64)["rgba8"]
    assert np.abs(img.astype(int) - two.astype(int)).max() <= 1
    # same-direction inner square + even-odd rule cuts the same hole
    enc = pm.Encoder(1 << 12)
    enc.begin_group(1)
    enc.fill_subpaths([outer, inner[::-1]], 0x203040ff, flags=pm.FILL_EVEN_ODD)
    enc.end_group()
    img2 = oracle.render(enc.bytes(), 96, 64, flags=PMO_FLAG_FILL_RULES)["rgba8"]
    assert np.array_equal(img2, img)
