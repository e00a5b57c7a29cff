"""ctypes bindings of the test infrastructure: the C oracle (oracle/pm_oracle.c) and the host
harness of the product's logic headers (tests/native/pm_host_harness.cpp)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "libpm_oracle.so")
HARNESS_PATH = os.path.join(ROOT, "tests", "native", "libpm_host_harness.so")
TILE_ITEM_DTYPE = np.dtype([("item", np.uint32), ("backdrop", np.int32), ("effect", np.uint32)])
CMD_DTYPE = np.dtype([("tag", np.uint32), ("body", np.uint32, 5)])

_libs = {}


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            raise ImportError("%s missing: run `python __graft_entry__.py` (build())" % path)
        _libs[path] = ctypes.CDLL(path)
    return _libs[path]


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _render(fn, scene, width, height, tile_y0, tile_y1, flags, threads, want_f32, want_items, want_rgba8=True):
    scene = np.ascontiguousarray(scene, np.uint8)
    nty, ntx = (height + 15) // 16, (width + 15) // 16
    if tile_y1 is None:
        tile_y1 = nty
    rows_px = min(tile_y1 * 16, height) - tile_y0 * 16
    n_tiles = (tile_y1 - tile_y0) * ntx
    out = {}
    rgba8 = np.zeros((rows_px, width, 4), np.uint8) if want_rgba8 else None
    f32 = np.zeros((rows_px, width, 4), np.float32) if want_f32 else None
    offsets = np.zeros(n_tiles + 1, np.uint32) if want_items else None
    solid = np.zeros(n_tiles, np.uint32) if want_items else None
    cap = max(4096, 8 * n_tiles) if want_items else 0
    while True:
        items = np.zeros(cap, TILE_ITEM_DTYPE) if want_items else None
        n = ctypes.c_size_t(0)
        args = [_vp(scene), ctypes.c_size_t(scene.size), ctypes.c_uint32(width), ctypes.c_uint32(height),
                ctypes.c_uint32(tile_y0), ctypes.c_uint32(tile_y1), ctypes.c_uint32(flags)]
        if threads is not None:
            args.append(ctypes.c_int(threads))
        args += [_vp(rgba8), ctypes.c_size_t(width * 4), _vp(f32), ctypes.c_size_t(width * 16),
                 _vp(offsets), _vp(items), ctypes.c_size_t(cap), ctypes.byref(n), _vp(solid)]
        rc = fn(*args)
        if rc != 0:
            raise ValueError("oracle rejected the scene (rc=%d)" % rc)
        if want_items and n.value > cap:
            cap = n.value
            continue
        break
    if want_rgba8:
        out["rgba8"] = rgba8
    if want_f32:
        out["rgba32f"] = f32
    if want_items:
        out["offsets"], out["items"], out["solid"] = offsets, items[:n.value], solid
    return out


def render(scene, width, height, tile_y0=0, tile_y1=None, flags=0, threads=0, f32=False, items=False, rgba8=True):
    """The oracle: scalar transliteration of the reference's tile loop (OpenMP over tile rows)."""
    fn = _load(ORACLE_PATH).pmo_render
    fn.restype = ctypes.c_int
    return _render(fn, scene, width, height, tile_y0, tile_y1, flags, threads, f32, items, rgba8)


def harness_render(scene, width, height, tile_y0=0, tile_y1=None, flags=0, f32=False, items=False, rgba8=True):
    """The product's binning/fill logic headers replayed on the CPU (test-only)."""
    fn = _load(HARNESS_PATH).pmh_render
    fn.restype = ctypes.c_int
    return _render(fn, scene, width, height, tile_y0, tile_y1, flags, None, f32, items, rgba8)


def tile_cmds(scene, tx, ty, flags=0):
    """The oracle's raw 24-byte Cmd stream of one tile and its solid colour."""
    lib = _load(ORACLE_PATH)
    scene = np.ascontiguousarray(scene, np.uint8)
    cap = 4096
    while True:
        cmds = np.zeros(cap, CMD_DTYPE)
        n = ctypes.c_size_t(0)
        solid = ctypes.c_uint32(0)
        rc = lib.pmo_tile_cmds(_vp(scene), ctypes.c_size_t(scene.size), ctypes.c_uint32(tx), ctypes.c_uint32(ty),
                               ctypes.c_uint32(flags), _vp(cmds), ctypes.c_size_t(cap), ctypes.byref(n), ctypes.byref(solid))
        if rc != 0:
            raise ValueError("oracle rejected the scene")
        if n.value > cap:
            cap = n.value
            continue
        return cmds[:n.value], solid.value


def max_threads():
    return _load(ORACLE_PATH).pmo_max_threads()


# ---------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own PietRender.metal compiled for the CPU (oracle/metal_shim/)
# ---------------------------------------------------------------------------------------------
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libpm_ref.so")
REF_HALF_PATH = os.path.join(ROOT, "oracle", "_ref", "libpm_ref_half.so")
REF_TILE_BUF = 4096
# bytes of a 24-byte Cmd that its tag defines (GenTypes.h:340-495); the rest is padding the reference never writes
CMD_FIELDS = {1: [], 9: [], 2: [(8, 16)], 3: [(8, 24)], 4: [(8, 24)], 5: [(4, 12)], 6: [(4, 12)], 7: [(4, 12)], 8: [(4, 8)]}


def have_ref(half=False):
    return os.path.exists(REF_HALF_PATH if half else REF_PATH)


def ref_render(scene, width, height, threads=0, half=False, want_cmds=False):
    """The reference's shaders (tileKernel, renderKernel, vertex/fragment composite) run by
    oracle/metal_shim/ref_driver.cpp.  Surfaces up to 4096 x 4096; n_cmds[tile] == -1 marks a tile whose
    list overflowed its 4096 bytes (the reference does not check) -- see `ref_trusted_tiles`."""
    lib = _load(REF_HALF_PATH if half else REF_PATH)
    scene = np.ascontiguousarray(scene, np.uint8)
    nty, ntx = (height + 15) // 16, (width + 15) // 16
    rgba8 = np.zeros((height, width, 4), np.uint8)
    f32 = np.zeros((height, width, 4), np.float32)
    solid = np.zeros(ntx * nty, np.uint32)
    n_cmds = np.zeros(ntx * nty, np.int32)
    cmds = np.zeros((ntx * nty, REF_TILE_BUF), np.uint8) if want_cmds else None
    rc = lib.pmref_render(_vp(scene), ctypes.c_size_t(scene.size), ctypes.c_uint32(width), ctypes.c_uint32(height), ctypes.c_int(threads),
                          _vp(rgba8), _vp(f32), _vp(solid), _vp(n_cmds), _vp(cmds))
    if rc != 0:
        raise ValueError("oracle/_ref rejected the frame (rc=%d): larger than 4096 x 4096?" % rc)
    out = {"rgba8": rgba8, "rgba32f": f32, "solid": solid, "n_cmds": n_cmds}
    if want_cmds:
        out["cmds"] = cmds
    return out


def ref_trusted_tiles(n_cmds):
    """Tiles whose command buffer is intact: neither overflowed nor overwritten by the left neighbour's overflow."""
    bad = n_cmds < 0
    excl = bad.copy()
    excl[1:] |= bad[:-1]
    return ~excl


def canonical_cmds(cmds):
    """A Cmd stream (CMD_DTYPE array or raw bytes) with every byte its tags do not define set to zero."""
    raw = np.ascontiguousarray(cmds).view(np.uint8).reshape(-1, 24)
    out = np.zeros_like(raw)
    out[:, 0:4] = raw[:, 0:4]
    tags = raw[:, 0:4].copy().view(np.uint32).reshape(-1)
    for tag, spans in CMD_FIELDS.items():
        rows = tags == tag
        for a, b in spans:
            out[rows, a:b] = raw[rows, a:b]
    return out
