"""The piet-style RenderContext facade (pm_context_*): recording on the host (CPU tests), finish() + rendering on the GPU."""
import numpy as np
import pytest


def drawing(pm, ctx):
    ring = pm.BezPath().move_to(20.3, 20.1).line_to(200.2, 24.4).line_to(196.1, 180.3).line_to(24.2, 176.2).close_path() \
        .move_to(70.4, 60.2).line_to(66.6, 140.1).line_to(150.3, 138.2).line_to(146.5, 62.3).close_path()   # inner wound the other way
    blob = pm.BezPath().move_to(30, 200).quad_to(120, 120, 220, 210).curve_to(180, 250, 90, 260, 30, 200)
    ctx.clear(0xf0f0e0ff)
    ctx.fill(ring, ctx.solid_brush(0x204080ff))
    ctx.save()
    ctx.transform([0.5, 0.1, -0.1, 0.5, 100.0, 20.0])
    ctx.fill_even_odd(blob, 0xc03020c0)
    ctx.stroke(blob, 0x000000ff, 6.0)
    ctx.restore()
    ctx.stroke(ring, 0x101010ff, 0.3)   # thinner than 0.7 px: the thin-stroke rule of src/lib.rs:353-362


def test_context_records_what_make_tiger_would_encode(pm):
    ctx = pm.RenderContext(None, 256, 256)
    drawing(pm, ctx)
    ps = ctx.path_set()
    # clear, compound ring fill, blob fill (even-odd), blob stroke, two ring strokes
    assert list(ps["tag"]) == [3, 3, 3, 4, 4, 4] and list(ps["flags"]) == [0, 0, 1, 0, 0, 0]
    # the ring: 3 segments + explicit close, bridge, 3 segments + close, bridge back = 3 + 1 + 1 + 3 + 1 + 1
    assert ps["first"][2] - ps["first"][1] == 10
    # stroke widths: 6.0 scaled by sqrt(|det|) of the transform; 0.3 -> 0.7 with the alpha scaled by sqrt(0.3 / 0.7)
    det = 0.5 * 0.5 + 0.1 * 0.1
    assert abs(ps["width"][3] - 6.0 * np.sqrt(det)) < 1e-5 and ps["width"][4] == np.float32(0.7)
    assert (ps["rgba"][4] & 0xff) == int(255 * np.sqrt(np.float32(0.3) / np.float32(0.7)))
    # the quad was raised to a cubic and transformed: its end point is the affine image of (220, 210)
    seg = ps["first"][2]
    assert ps["verb"][seg] == 1 and np.allclose(ps["ctrl"][seg][4:], [0.5 * 220 - 0.1 * 210 + 100, 0.1 * 220 + 0.5 * 210 + 20])
    with pytest.raises(pm.PietMetalError):
        ctx.restore()                       # nothing saved
    with pytest.raises(pm.PietMetalError):
        ctx.finish()                        # no renderer behind this context


@pytest.mark.gpu
def test_context_finish_renders_like_the_oracle(pm, oracle):
    r = pm.PietRenderer(device=0, flags=pm.FLAG_FILL_RULES)
    try:
        r.drawable_size_will_change(256, 256)
        ctx = pm.RenderContext(r)
        drawing(pm, ctx)
        ctx.finish()
        assert ctx.item_count() == 0        # finish() starts a new drawing
        r.draw()
        img = r.read_rgba8()
        scene = r.read_scene()              # what the device encoded
        assert pm.validate_scene(scene) == 0
        ref = oracle.render(scene, 256, 256, flags=pm.FLAG_FILL_RULES)["rgba8"]
        assert np.abs(img.astype(int) - ref.astype(int)).max() <= 1
        assert (img[100, 110, :3] == [0xf0, 0xf0, 0xe0]).all()          # inside the ring's hole: the cleared background
        assert (img[40, 40, :3] == [0x20, 0x40, 0x80]).all()            # in the ring
        assert (img[250, 250, :3] == [0xf0, 0xf0, 0xe0]).all()
        ctx.finish()                        # an empty drawing is the white background
        r.draw()
        assert (r.read_rgba8() == 255).all()
    finally:
        r.close()
