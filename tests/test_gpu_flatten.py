"""Flattening + scene encoding on the device (pm_renderer_set_scene_paths) against the CPU feed.

The reference flattens and encodes on the CPU (src/flatten.rs:10-47, src/lib.rs:195-240); pm_feed.cpp restates that
(pm_flatten_svg_path + Encoder).  The device path must produce the same encoded scene: same item and point counts, same
bounding boxes and items, points equal to the last bit -- it evaluates the same f64 expressions in the same order --
except where CUDA's pow() and glibc's differ in the ulp that decides a subdivision count (reported, and bounded)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_paths(pm, seed, n_paths, size):
    """Random subpaths of lines and cubics, as a PathSet for the device and as the scene the CPU feed encodes from the
    same numbers (through SVG path data: repr() of a double round-trips through strtod)."""
    rng = np.random.default_rng(seed)
    ps = pm.PathSet()
    cpu = []
    for _ in range(n_paths):
        stroke = rng.random() < 0.3
        rgba = (int(rng.integers(0, 1 << 24)) << 8) | int(rng.choice([255, 128, 40]))
        width = float(rng.choice([0.7, 1.5, 4.0])) if stroke else 0.0
        x, y = rng.uniform(0, size, 2)
        ps.begin(x, y, 4 if stroke else 3, rgba, width)
        d = "M%r %r" % (float(x), float(y))
        for _ in range(int(rng.integers(0, 9))):
            r = float(rng.choice([4.0, 30.0, 300.0]))
            if rng.random() < 0.35:
                x, y = x + rng.uniform(-r, r), y + rng.uniform(-r, r)
                ps.line_to(x, y)
                d += "L%r %r" % (float(x), float(y))
            else:
                c = [x + rng.uniform(-r, r), y + rng.uniform(-r, r), x + rng.uniform(-r, r), y + rng.uniform(-r, r)]
                x, y = x + rng.uniform(-r, r), y + rng.uniform(-r, r)
                ps.curve_to(c[0], c[1], c[2], c[3], x, y)
                d += "C%r %r %r %r %r %r" % (float(c[0]), float(c[1]), float(c[2]), float(c[3]), float(x), float(y))
        cpu.append((d, stroke, rgba, width))
    return ps, cpu


def cpu_scene(pm, cpu, scale, tolerance):
    subs = []
    for d, stroke, rgba, width in cpu:
        sp = pm.flatten_svg_path(d, scale, tolerance)
        assert len(sp) == 1
        subs.append((sp[0], stroke, rgba, width))
    enc = pm.Encoder(64 + sum(40 + 8 * len(s[0]) for s in subs))
    enc.begin_group(len(subs))
    for pts, stroke, rgba, width in subs:
        if stroke:
            enc.polyline(pts, rgba, width)
        else:
            enc.fill(pts, rgba)
    enc.end_group()
    return enc.bytes()


@pytest.mark.parametrize("seed,n_paths,scale", [(1, 50, 1.0), (2, 2000, 1.0), (3, 2000, 3.7), (4, 20000, 0.5)])
def test_device_flattening_matches_cpu_feed(pm, seed, n_paths, scale):
    size = 2048
    ps, cpu = random_paths(pm, seed, n_paths, size / scale)
    want = cpu_scene(pm, cpu, scale, 0.1)
    r = pm.PietRenderer(device=0)
    try:
        r.drawable_size_will_change(size, size)
        r.set_scene_paths(ps, scale=scale, tolerance=0.1)
        got = r.read_scene()
        n = n_paths
        assert got[:8].tobytes() == want[:8].tobytes()                      # header
        gi, wi = got[8 + 8 * n:8 + 40 * n].view(np.uint32).reshape(n, 8), want[8 + 8 * n:8 + 40 * n].view(np.uint32).reshape(n, 8)
        same_counts = gi[:, 3] == wi[:, 3]
        # pow() may differ in its last bit between CUDA and glibc; that can move a subdivision count only on a knife edge
        assert same_counts.mean() > 0.999, "point counts differ for %d of %d subpaths" % ((~same_counts).sum(), n)
        if same_counts.all():
            assert got.size == want.size
            assert np.array_equal(got, want), "encoded scene differs from the CPU feed's"
            r.draw()
            img = r.read_rgba8()
            r.init_scene(want)
            r.draw()
            assert np.array_equal(img, r.read_rgba8())
    finally:
        r.close()


def test_device_flattening_tiger_like_and_errors(pm):
    r = pm.PietRenderer(device=0)
    try:
        r.drawable_size_will_change(256, 256)
        ps = pm.PathSet()
        ps.begin(10.0, 10.0, 3, 0x3366ccff)          # a subpath of a MoveTo only
        ps.begin(20.0, 20.0, 3, 0x112233ff, flags=1)
        ps.line_to(200.0, 30.0)
        ps.curve_to(220.0, 100.0, 120.0, 220.0, 30.0, 200.0)
        r.set_scene_paths(ps)
        scene = r.read_scene()
        assert pm.validate_scene(scene) == 0
        items = scene[8 + 16:8 + 16 + 64].view(np.uint32).reshape(2, 8)
        assert items[0, 0] == 3 and items[0, 3] == 1 and items[1, 1] == 1 and items[1, 3] >= 3
        r.draw()
        assert (r.read_rgba8()[100, 100, :3] != 255).any()
        bad = pm.PathSet()
        bad.begin(0.0, 0.0, 7, 0xff)                  # no such item tag
        with pytest.raises(pm.PietMetalError):
            r.set_scene_paths(bad)
    finally:
        r.close()
