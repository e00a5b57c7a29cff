"""Pins the oracle (oracle/pm_oracle.c) to the reference itself.

oracle/_ref/libpm_ref.so is the reference's UNMODIFIED TestApp/PietRender.metal (+ GenTypes.h,
PietShaderTypes.h) compiled with g++ through the metal_stdlib stand-in of oracle/metal_shim/ and
driven the way TestApp/PietRenderer.m:59-99 drives it.  Two anchors:
  * live: wherever the library exists (built in the dev container from /root/reference; it travels to
    the GPU box as a built file), the oracle's per-tile 24-byte command streams and solid colours
    must be bit-equal to tileKernel's and its pixels within 1e-6 of renderKernel's + composite
    (they are equal to the last bit: `half` is fp32 in that build and the arithmetic is the same);
  * golden: tests/golden/ref_vectors.json holds digests of what the reference produced here
    (tools/make_ref_golden.py); the oracle has to reproduce them even without the library.
The reference's hard limits apply to it, not to the oracle: 4096 x 4096 pixels, 170 commands per
tile with no overflow check -- tiles that overflow (one on the 1024^2 and 2048^2 tiger, the crowded first
rows of the glyph scene) and their right-hand neighbours are skipped."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")


def pin_cases(pm):
    """(name, scene bytes, width, height) of every pinned case -- shared with tools/make_ref_golden.py."""
    cases = [
        ("rect1_16", pm.build_scene(pm.SCENE_RECT1, 16, 16), 16, 16),
        ("path_test_512x832", pm.build_scene(pm.SCENE_PATH_TEST, 512, 832), 512, 832),       # src/lib.rs:273-284
        ("cardioid_2048x1536", pm.build_scene(pm.SCENE_CARDIOID, 2048, 1536), 2048, 1536),   # src/lib.rs:257-270
        ("tiger_scale8_1024x768", pm.build_scene(pm.SCENE_TIGER, 1024, 768, scale=8.0), 1024, 768),  # the reference's own window
        ("tiger_1024", pm.build_scene(pm.SCENE_TIGER, 1024, 1024), 1024, 1024),
        ("tiger_2048", pm.build_scene(pm.SCENE_TIGER, 2048, 2048), 2048, 2048),
        ("tiger_4096", pm.build_scene(pm.SCENE_TIGER, 4096, 4096), 4096, 4096),             # the reference's largest surface
        ("rand_bezier_400_1024", pm.build_scene(pm.SCENE_RAND_BEZIER, 1024, 1024, count=400), 1024, 1024),
        ("glyphs_3000_1024", pm.build_scene(pm.SCENE_GLYPHS, 1024, 1024, count=3000), 1024, 1024),
    ]
    for seed in range(40):
        scene, w, h, flags = scenes.fuzz_case(pm, seed)
        if flags == 0:  # (the FIX_POLY_PRECULL variant is not the reference's behaviour)
            cases.append(("fuzz_%02d" % seed, scene, w, h))
    return cases


def oracle_tile_streams(oracle, scene, width, height, tiles):
    ntx = (width + 15) // 16
    out = {}
    for t in tiles:
        cmds, solid = oracle.tile_cmds(scene, int(t) % ntx, int(t) // ntx)
        if cmds["tag"][0] == 9:  # Bail at tileBegin (metal:145-147): nothing behind it is ever read
            cmds = cmds[:1]
        out[int(t)] = (oracle.canonical_cmds(cmds), solid)
    return out


def digest_case(oracle, render, scene, width, height, trusted, streams):
    """Digests over the trusted tiles only: pixels (fp32 and RGBA8), solid colours, command streams."""
    nty, ntx = (height + 15) // 16, (width + 15) // 16
    mask = np.repeat(np.repeat(trusted.reshape(nty, ntx), 16, 0), 16, 1)[:height, :width]
    h = {}
    h["rgba8"] = hashlib.sha256(np.where(mask[..., None], render["rgba8"], 0).tobytes()).hexdigest()
    h["rgba32f"] = hashlib.sha256(np.where(mask[..., None], render["rgba32f"], 0).astype(np.float32).tobytes()).hexdigest()
    h["solid"] = hashlib.sha256(np.where(trusted, render["solid"], 0).astype(np.uint32).tobytes()).hexdigest()
    hc = hashlib.sha256()
    for t in sorted(streams):
        hc.update(np.uint32(t).tobytes())
        hc.update(streams[t][0].tobytes())
    h["cmds"] = hc.hexdigest()
    return h


def stream_tiles(trusted, limit=600):
    """The tiles whose command streams are compared: all of them on small frames, a seeded sample otherwise."""
    idx = np.flatnonzero(trusted)
    if len(idx) > limit:
        idx = np.sort(np.random.default_rng(12345).choice(idx, limit, replace=False))
    return idx


def test_oracle_matches_reference_live(pm, oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libpm_ref.so not built (needs /root/reference)")
    for name, scene, w, h in pin_cases(pm):
        ref = oracle.ref_render(scene, w, h, want_cmds=True)
        trusted = oracle.ref_trusted_tiles(ref["n_cmds"])
        assert trusted.mean() > 0.9, name
        ours = oracle.render(scene, w, h, f32=True, items=True)
        nty, ntx = (h + 15) // 16, (w + 15) // 16
        mask = np.repeat(np.repeat(trusted.reshape(nty, ntx), 16, 0), 16, 1)[:h, :w]
        assert np.array_equal(ours["solid"][trusted], ref["solid"][trusted]), name
        d32 = np.abs(ours["rgba32f"] - ref["rgba32f"])[mask]
        assert d32.max() <= 1e-6, (name, d32.max())
        assert np.array_equal(ours["rgba8"][mask], ref["rgba8"][mask]), name
        for t, (stream, solid) in oracle_tile_streams(oracle, scene, w, h, stream_tiles(trusted)).items():
            n = int(ref["n_cmds"][t])  # (both lists include their End / Bail)
            assert np.array_equal(stream, oracle.canonical_cmds(ref["cmds"][t][:n * 24])), (name, t)
            assert solid == ref["solid"][t], (name, t)


def test_oracle_matches_reference_golden(pm, oracle):
    with open(GOLDEN) as f:
        golden = json.load(f)
    cases = {name: (scene, w, h) for name, scene, w, h in pin_cases(pm)}
    assert set(golden["cases"]) == set(cases)
    for name, g in golden["cases"].items():
        scene, w, h = cases[name]
        assert hashlib.sha256(scene.tobytes()).hexdigest() == g["scene_sha256"], "%s: the feed no longer encodes the pinned scene" % name
        nty, ntx = (h + 15) // 16, (w + 15) // 16
        trusted = np.ones(ntx * nty, bool)
        trusted[np.array(g["untrusted_tiles"], np.int64)] = False
        ours = oracle.render(scene, w, h, f32=True, items=True)
        streams = oracle_tile_streams(oracle, scene, w, h, g["stream_tiles"])
        d = digest_case(oracle, ours, scene, w, h, trusted, streams)
        for k in ("rgba8", "rgba32f", "solid", "cmds"):
            assert d[k] == g[k], (name, k)


def test_half_build_documents_the_precision_gap(pm, oracle):
    """The reference accumulates colour, alpha and signedArea in `half` (metal:470-472, :502, :526, :537);
    the task's contract is fp32.  With half = _Float16 the reference's frame is a few LSB away."""
    if not oracle.have_ref(half=True):
        pytest.skip("oracle/_ref/libpm_ref_half.so not built")
    scene = pm.build_scene(pm.SCENE_TIGER, 512, 512)
    a = oracle.ref_render(scene, 512, 512)
    b = oracle.ref_render(scene, 512, 512, half=True)
    trusted = oracle.ref_trusted_tiles(a["n_cmds"])
    mask = np.repeat(np.repeat(trusted.reshape(32, 32), 16, 0), 16, 1)
    d = np.abs(a["rgba8"].astype(int) - b["rgba8"].astype(int))[mask]
    assert np.array_equal(a["solid"], b["solid"])
    assert 1 <= d.max() <= 8
