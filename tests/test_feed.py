"""Feed behaviour: SVG path data, flattening, colours, the reference's test scenes (src/lib.rs, src/flatten.rs)."""
import numpy as np
import pytest


def test_parse_color(pm):   # src/lib.rs:375-385
    assert pm.parse_color("#FFF") == 0xFFFFFFFF
    assert pm.parse_color("#cc7226") == 0xCC7226FF
    assert pm.parse_color("#1a2") == 0x11AA22FF
    assert pm.parse_color("none") == 0xFF00FF80


def test_svg_path_commands_and_compact_numbers(pm):
    sub = pm.flatten_svg_path("M10 10L20 10l0 5h-5v5H10z")
    assert len(sub) == 1
    assert sub[0].tolist() == [[10, 10], [20, 10], [20, 15], [15, 15], [15, 20], [10, 20]]
    # compact forms used by the tiger: ".039.744", "0-3.551", implicit lineto after moveto
    sub = pm.flatten_svg_path("M1 2.5.5 3-1-2")
    assert sub[0].tolist() == [[1, 2.5], [0.5, 3], [-1, -2]]
    # z then a relative moveto starts from the subpath start (SVG 1.1, 8.3.3)
    sub = pm.flatten_svg_path("M10 10l5 0l0 5zm1 1l2 0")
    assert len(sub) == 2 and sub[1].tolist() == [[11, 11], [13, 11]]
    with pytest.raises(pm.PietMetalError):
        pm.flatten_svg_path("M10 10 L")


def test_flatten_cubic_follows_to_quads_rule(pm):
    """flatten.rs:28-37 with kurbo's CubicBez::to_quads(tol*1e-2): n = ceil((err / (432 acc^2))^(1/6))
    uniform parameter steps, end points of the pieces only."""
    p0, p1, p2, p3 = np.array([0.0, 0.0]), np.array([0.0, 100.0]), np.array([100.0, 100.0]), np.array([100.0, 0.0])
    sub = pm.flatten_svg_path("M0 0C0 100 100 100 100 0", tolerance=0.1)
    acc = 0.1 * 1e-2
    err = np.sum(((3 * p2 - p3) - (3 * p1 - p0)) ** 2)
    n = int(max(1, np.ceil((err / (432 * acc * acc)) ** (1 / 6))))
    assert len(sub[0]) == n + 1
    t = np.arange(1, n + 1) / n
    mt = 1 - t
    want = (np.outer(mt ** 3, p0) + np.outer(3 * mt * mt * t, p1) + np.outer(3 * mt * t * t, p2) + np.outer(t ** 3, p3))
    assert np.allclose(sub[0][1:], want, atol=1e-9)
    assert sub[0][-1].tolist() == [100.0, 0.0]
    # scale multiplies the points before flattening (kurbo::Affine::scale * path, lib.rs:314)
    sub2 = pm.flatten_svg_path("M0 0C0 100 100 100 100 0", scale=2.0)
    assert sub2[0][-1].tolist() == [200.0, 0.0] and len(sub2[0]) > len(sub[0])
    # QuadTo is ignored by flatten_path (flatten.rs:40)
    assert pm.flatten_svg_path("M0 0Q5 5 10 0L20 0")[0].tolist() == [[0, 0], [20, 0]]


def test_arc_ends_where_it_should(pm):
    sub = pm.flatten_svg_path("M10 0a10 10 0 0 1-20 0")
    assert np.allclose(sub[0][-1], [-10.0, 0.0], atol=1e-9)
    r = np.hypot(sub[0][:, 0], sub[0][:, 1])
    assert np.abs(r - 10.0).max() < 0.02  # cubic approximation of a half circle, tolerance 0.1


def test_tiger_item_census(pm):
    """Appendix A of SURVEY.md: 226 fill subpaths + 78 stroke subpaths = 304 items."""
    scene = pm.build_scene(pm.SCENE_TIGER, 1024, 1024)
    n = int(scene[:4].view(np.uint32)[0])
    items_ix = int(scene[4:8].view(np.uint32)[0])
    tags = scene[items_ix:items_ix + 32 * n].view(np.uint32).reshape(n, 8)[:, 0]
    assert n == 304 and int((tags == 3).sum()) == 226 and int((tags == 4).sum()) == 78
    assert pm.validate_scene(scene) == 0
    bbox = scene[8:8 + 8 * n].view(np.uint16).reshape(n, 4)
    assert bbox[:, 2].max() <= 1040 and bbox[:, 3].max() <= 1040  # the artwork slightly overshoots its 200x200 viewBox


def test_init_test_scene_is_the_tiger_at_scale_8(pm):
    """init_test_scene (include/piet_metal.h:3) == make_tiger with scale 8.0 (lib.rs:287,369-373)."""
    buf = pm.init_test_scene(4 << 20)
    want = pm.build_scene(pm.SCENE_TIGER, 1600, 1600, scale=8.0)
    assert np.array_equal(buf[:want.size], want)
    assert not buf[want.size:].any()
    small = pm.init_test_scene(1024)   # too small: nothing past the buffer, empty group
    assert small[:4].view(np.uint32)[0] == 0


def test_thin_stroke_fudge(pm):
    """encode_path_stroke (lib.rs:353-362): widths below 0.7 px are clamped and alpha scaled by sqrt."""
    text = b"path - #000 .05 M0 0L100 0\n"
    import ctypes
    lib = pm._lib()
    n = lib.pm_scene_from_pathlist(text, len(text), 8.0, None, 0)
    buf = np.zeros(n, np.uint8)
    assert lib.pm_scene_from_pathlist(text, len(text), 8.0, buf.ctypes.data_as(ctypes.c_void_p), n) == n
    item = buf[16:48].view(np.uint32)
    assert item[0] == 4
    width = buf[16 + 8:16 + 12].view(np.float32)[0]
    assert width == np.float32(0.7)
    alpha = int(np.float32(255.0) * np.sqrt(np.float32(0.05 * 8.0) / np.float32(0.7)))
    assert item[1] >> 24 == alpha


def test_synthetic_scenes_are_deterministic(pm):
    a = pm.build_scene(pm.SCENE_RAND_BEZIER, 512, 512, count=50)
    b = pm.build_scene(pm.SCENE_RAND_BEZIER, 512, 512, count=50)
    c = pm.build_scene(pm.SCENE_RAND_BEZIER, 512, 512, count=50, seed=123)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    g = pm.build_scene(pm.SCENE_GLYPHS, 512, 512, count=300)
    assert int(g[:4].view(np.uint32)[0]) == 300 and pm.validate_scene(g) == 0


def test_png_and_ppm_egress_round_trip(pm, oracle, tmp_path):
    """pm_write_png / pm_write_ppm: decode what was written (PNG by hand: signature, chunk CRCs, zlib stream)."""
    import struct
    import zlib
    w, h = 70, 37
    scene = pm.build_scene(pm.SCENE_TIGER, 128, 128)
    img = oracle.render(scene, 128, 128)["rgba8"][:h, :w].copy()
    png, ppm = str(tmp_path / "t.png"), str(tmp_path / "t.ppm")
    pm.write_image(png, img)
    pm.write_image(ppm, img)
    data = open(png, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + body)
        chunks[typ] = body
        pos += 12 + n
    assert struct.unpack(">IIBBBBB", chunks[b"IHDR"]) == (w, h, 8, 6, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(h, w, 4), img)
    p = open(ppm, "rb").read()
    header = b"P6\n%d %d\n255\n" % (w, h)
    assert p.startswith(header) and np.array_equal(np.frombuffer(p[len(header):], np.uint8).reshape(h, w, 3), img[:, :, :3])
    # a frame larger than one stored block (65535 bytes)
    big = np.arange(300 * 200 * 4, dtype=np.uint32).astype(np.uint8).reshape(200, 300, 4)
    pm.write_image(png, big)
    d = open(png, "rb").read()
    i = d.index(b"IDAT")
    n = struct.unpack(">I", d[i - 4:i])[0]
    assert np.array_equal(np.frombuffer(zlib.decompress(d[i + 4:i + 4 + n]), np.uint8).reshape(200, 1201)[:, 1:].reshape(200, 300, 4), big)
