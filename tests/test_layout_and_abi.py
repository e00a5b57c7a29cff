"""Wire format of the scene buffer and the C-ABI surface -- no GPU needed."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest


def test_path_test_scene_golden_bytes(pm):
    """make_path_test (src/lib.rs:273-284): 8 B header + 8 B bbox + 32 B item + 24 B points, hand-derived."""
    scene = pm.build_scene(pm.SCENE_PATH_TEST, 320, 816)
    assert scene.size == 72
    want = struct.pack("<II", 1, 16)                       # n_items, items_ix = 8 + 8*1
    want += struct.pack("<4H", 10, 10, 300, 800)           # bbox floor/floor/ceil/ceil
    want += struct.pack("<5I", 3, 0, 0xE0800000, 3, 48)    # Fill, flags, rgba = 0x000080e0.to_be(), n_points, points_ix
    want += b"\0" * 12
    want += struct.pack("<6f", 10, 10, 15, 800, 300, 500)
    assert scene.tobytes() == want


def test_rect1_scene_layout(pm):
    """BASELINE config 1: 80-byte scene (8 hdr + 8 bbox + 32 item + 32 points)."""
    scene = pm.build_scene(pm.SCENE_RECT1, 16, 16, rect=(3.25, 2.5, 12.75, 13.5))
    assert scene.size == 80
    n, items_ix = struct.unpack_from("<II", scene, 0)
    assert (n, items_ix) == (1, 16)
    assert struct.unpack_from("<4H", scene, 8) == (3, 2, 13, 14)
    tag, flags, rgba, npts, pix = struct.unpack_from("<5I", scene, 16)
    assert (tag, npts, pix) == (3, 4, 48)
    assert rgba == 0xFFCC6633  # 0x3366ccff stored big-endian: bytes R,G,B,A in memory
    assert struct.unpack_from("<8f", scene, 48) == (3.25, 2.5, 12.75, 2.5, 12.75, 13.5, 3.25, 13.5)


def test_encoder_mirrors_reference_api(pm):
    """Encoder::{begin_group, circle, stroke_line, fill, polyline, end_group} (src/lib.rs:132-222)."""
    enc = pm.Encoder(4096)
    enc.begin_group(4)
    enc.circle(100.0, 50.0, 8.0)
    enc.stroke_line((1.5, 2.5), (30.0, 40.0), 2.0, 0x000080E0)
    enc.fill(np.array([(0.0, 0.0), (10.0, 0.0), (10.0, 10.0)]), 0x112233FF)
    enc.polyline(np.array([(5.0, 5.0), (6.0, 9.0)]), 0xAABBCC80, 3.0)
    enc.end_group()
    s = enc.bytes()
    n, items_ix = struct.unpack_from("<II", s, 0)
    assert n == 4 and items_ix == 8 + 8 * 4
    bboxes = np.frombuffer(s, np.uint16, 16, 8).reshape(4, 4)
    assert bboxes[0].tolist() == [92, 42, 108, 58]           # circle bbox
    assert bboxes[1].tolist() == [0, 1, 31, 41]              # line bbox inflated by width/2
    assert bboxes[3].tolist() == [3, 3, 8, 11]               # polyline bbox inflated by width/2 = 1.5
    items = np.frombuffer(s, np.uint32, 32, items_ix).reshape(4, 8)
    assert items[:, 0].tolist() == [1, 2, 3, 4]
    assert items[1, 2] == 0xE0800000 and struct.unpack_from("<f", s, items_ix + 32 + 12)[0] == 2.0
    assert items[2, 3] == 3 and items[3, 3] == 2
    assert items[3, 1] == 0x80CCBBAA                          # PietStrokePolyLine: rgba at +4
    assert s.size == items_ix + 4 * 32 + 8 * 5
    assert pm.validate_scene(s) == 0
    with pytest.raises(pm.PietMetalError):                   # assert!(group_ix < group_count), lib.rs:152
        enc.circle(0, 0, 1)


def test_encoder_overflow_is_reported_not_written(pm):
    enc = pm.Encoder(64)
    with pytest.raises(pm.PietMetalError) as e:
        enc.begin_group(10)
        enc.fill(np.zeros((50, 2)), 0xFF)
    assert e.value.status == pm.PM_ERR_BUFFER_TOO_SMALL


def test_validate_rejects_out_of_range_refs(pm):
    scene = pm.build_scene(pm.SCENE_PATH_TEST, 320, 816).copy()
    assert pm.validate_scene(scene) == 0
    bad = scene.copy(); bad[32:36].view(np.uint32)[0] = 1 << 20   # points_ix
    assert pm.validate_scene(bad) == pm.PM_ERR_SCENE_MALFORMED
    bad = scene.copy(); bad[28:32].view(np.uint32)[0] = 0          # n_points = 0
    assert pm.validate_scene(bad) == pm.PM_ERR_SCENE_MALFORMED
    bad = scene.copy(); bad[0:4].view(np.uint32)[0] = 1000         # n_items beyond the buffer
    assert pm.validate_scene(bad) == pm.PM_ERR_SCENE_MALFORMED
    # refs must be 8-byte aligned (the kernels read points and Line end points with 64-bit loads; the reference's
    # encoder only ever produces such refs, src/lib.rs:132-163, :224-240): 4 mod 8 is rejected, not faulted on
    pix = int(scene[32:36].view(np.uint32)[0])
    bad = np.concatenate([scene, np.zeros(8, np.uint8)]); bad[32:36].view(np.uint32)[0] = pix + 4
    assert pm.validate_scene(bad) == pm.PM_ERR_SCENE_MALFORMED


def test_library_exports_every_declared_symbol(pm):
    """Every function declared in include/piet_metal_b200.h is exported by the shared library."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "piet_metal_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b((?:pm_|init_test_scene)\w*)\s*\(", header))
    declared -= {"pm_status", "pm_scene_kind"}
    lib = ctypes.CDLL(pm.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, "declared but not exported: %s" % missing
    assert set(pm.EXPORTS) == declared
    assert re.match(r"\d+\.\d+\.\d+$", pm.version())


def test_renderer_fails_loudly_without_a_gpu(pm):
    """No CPU fallback: without a CUDA device the renderer refuses to exist."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pm.PietMetalError) as e:
        pm.PietRenderer(device=0)
    assert e.value.status == pm.PM_ERR_NO_DEVICE
