"""piet_metal_b200 -- Python host-side mirror of the reference's operator interface.

The product is the C-ABI shared library ``libpiet_metal_b200.so`` (``include/piet_metal_b200.h``);
this module is the thin ctypes binding the tests and ``bench.py`` drive it through.  Names follow the
reference: ``Encoder`` mirrors the Rust ``Encoder`` (src/lib.rs:79-254) and ``PietRenderer`` mirrors
the Objective-C ``PietRenderer`` (TestApp/PietRenderer.m: ``initWithMetalKitView:`` :23,
``mtkView:drawableSizeWillChange:`` :105, ``initScene`` :203, ``drawInMTKView:`` :59).

There is no CPU rendering path: if the library is missing, or there is no sm_100 GPU, the renderer
raises instead of falling back.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PM_LIB") or os.path.join(_HERE, "libpiet_metal_b200.so")  # PM_LIB: A/B builds of the same library

PM_OK = 0
PM_ERR_NO_DEVICE = -2
PM_ERR_SCENE_MALFORMED = -4
PM_ERR_BUFFER_TOO_SMALL = -5
PM_ERR_STATE = -6

SCENE_RECT1, SCENE_PATH_TEST, SCENE_CARDIOID, SCENE_TIGER, SCENE_RAND_BEZIER, SCENE_GLYPHS = range(6)
FLAG_FIX_POLY_PRECULL = 1
FLAG_EXACT_SRGB = 2
FLAG_FILL_RULES = 4
FILL_NONZERO, FILL_EVEN_ODD = 0, 1
SCENE_OPT_COMPOUND_FILLS, SCENE_OPT_EVEN_ODD = 1, 2


class PietMetalError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        lib = _lib()
        msg = lib.pm_strerror(status).decode()
        last = lib.pm_last_error().decode()
        super().__init__("%s: %s (%d)%s" % (where, msg, status, (" -- " + last) if last else ""))


class SceneDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32),
                ("count", ctypes.c_uint32), ("seed", ctypes.c_uint64), ("scale", ctypes.c_double),
                ("rect", ctypes.c_double * 4), ("rgba", ctypes.c_uint32), ("options", ctypes.c_uint32)]


class Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("flags", ctypes.c_uint32), ("scratch_bytes", ctypes.c_uint64)]


class FrameStats(ctypes.Structure):
    _fields_ = [("ms_total", ctypes.c_float), ("ms_bin", ctypes.c_float), ("ms_fine", ctypes.c_float),
                ("frames", ctypes.c_uint32), ("ms_total_sum", ctypes.c_float), ("ms_bin_sum", ctypes.c_float),
                ("ms_fine_sum", ctypes.c_float), ("n_tiles", ctypes.c_uint32), ("n_overflow_records", ctypes.c_uint32),
                ("n_complex_tiles", ctypes.c_uint32), ("n_launches", ctypes.c_uint32), ("retries", ctypes.c_uint32),
                ("ms_heavy", ctypes.c_float), ("ms_heavy_sum", ctypes.c_float), ("n_heavy_tiles", ctypes.c_uint32),
                ("ms_plan", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PathSetC(ctypes.Structure):
    _fields_ = [("n_subpaths", ctypes.c_uint32), ("n_segments", ctypes.c_uint32), ("first_segment", ctypes.c_void_p),
                ("start", ctypes.c_void_p), ("verb", ctypes.c_void_p), ("ctrl", ctypes.c_void_p), ("tag", ctypes.c_void_p),
                ("rgba", ctypes.c_void_p), ("width", ctypes.c_void_p), ("flags", ctypes.c_void_p)]


TILE_ITEM_DTYPE = np.dtype([("item", np.uint32), ("backdrop", np.int32), ("effect", np.uint32)])

# every entry point declared in include/piet_metal_b200.h
EXPORTS = [
    "pm_strerror", "pm_last_error", "pm_version", "init_test_scene",
    "pm_encoder_new", "pm_encoder_begin_group", "pm_encoder_end_group", "pm_encoder_circle",
    "pm_encoder_stroke_line", "pm_encoder_fill", "pm_encoder_polyline", "pm_encoder_bytes", "pm_encoder_free",
    "pm_encoder_fill_rule", "pm_encoder_fill_subpaths",
    "pm_flatten_svg_path", "pm_parse_color", "pm_scene_build", "pm_scene_from_pathlist", "pm_scene_validate",
    "pm_scene_row_costs", "pm_balance_strips", "pm_write_ppm", "pm_write_png",
    "pm_renderer_create", "pm_renderer_destroy", "pm_renderer_resize", "pm_renderer_set_strip",
    "pm_renderer_set_scene", "pm_renderer_set_scene_device", "pm_renderer_render", "pm_renderer_set_frame_events", "pm_renderer_sync",
    "pm_renderer_read_rgba8", "pm_renderer_render_host", "pm_renderer_framebuffer", "pm_renderer_stream",
    "pm_renderer_read_rgba32f", "pm_renderer_read_tile_items", "pm_host_alloc", "pm_host_free",
    "pm_renderer_set_scene_paths", "pm_renderer_read_scene",
    "pm_context_new", "pm_context_free", "pm_context_save", "pm_context_restore", "pm_context_transform", "pm_context_clear",
    "pm_context_fill", "pm_context_fill_even_odd", "pm_context_stroke", "pm_context_item_count", "pm_context_path_set", "pm_context_finish",
    "pm_group_create", "pm_group_destroy", "pm_group_size", "pm_group_member", "pm_group_resize", "pm_group_set_scene",
    "pm_group_strip_bounds", "pm_group_set_frame_events", "pm_group_render", "pm_group_sync", "pm_group_read_rgba8",
    "pm_group_gather_device", "pm_group_nccl_version",
]

_LIB = None


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C %s` (or __graft_entry__.build()); "
                          "there is no fallback implementation" % (LIB_PATH, _HERE))
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u32, u64, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64
    dbl, flt, cint = ctypes.c_double, ctypes.c_float, ctypes.c_int
    sig = {
        "pm_strerror": (ctypes.c_char_p, [cint]),
        "pm_last_error": (ctypes.c_char_p, []),
        "pm_version": (ctypes.c_char_p, []),
        "init_test_scene": (None, [vp, ctypes.c_ssize_t]),
        "pm_encoder_new": (cint, [ctypes.POINTER(vp), vp, sz]),
        "pm_encoder_begin_group": (cint, [vp, u32]),
        "pm_encoder_end_group": (cint, [vp]),
        "pm_encoder_circle": (cint, [vp, dbl, dbl, dbl]),
        "pm_encoder_stroke_line": (cint, [vp, dbl, dbl, dbl, dbl, flt, u32]),
        "pm_encoder_fill": (cint, [vp, vp, u32, u32]),
        "pm_encoder_polyline": (cint, [vp, vp, u32, u32, flt]),
        "pm_encoder_fill_rule": (cint, [vp, vp, u32, u32, u32]),
        "pm_encoder_fill_subpaths": (cint, [vp, vp, vp, u32, u32, u32]),
        "pm_encoder_bytes": (sz, [vp]),
        "pm_encoder_free": (None, [vp]),
        "pm_flatten_svg_path": (i64, [ctypes.c_char_p, dbl, dbl, vp, sz, vp, sz, ctypes.POINTER(sz)]),
        "pm_parse_color": (u32, [ctypes.c_char_p]),
        "pm_scene_build": (i64, [ctypes.POINTER(SceneDesc), vp, sz]),
        "pm_scene_from_pathlist": (i64, [ctypes.c_char_p, sz, dbl, vp, sz]),
        "pm_scene_validate": (cint, [vp, sz]),
        "pm_scene_row_costs": (cint, [vp, sz, u32, u32, vp, sz]),
        "pm_balance_strips": (cint, [vp, u32, u32, vp]),
        "pm_write_ppm": (cint, [ctypes.c_char_p, vp, u32, u32, sz]),
        "pm_write_png": (cint, [ctypes.c_char_p, vp, u32, u32, sz]),
        "pm_renderer_create": (cint, [ctypes.POINTER(vp), ctypes.POINTER(Config)]),
        "pm_renderer_destroy": (None, [vp]),
        "pm_renderer_resize": (cint, [vp, u32, u32]),
        "pm_renderer_set_strip": (cint, [vp, u32, u32]),
        "pm_renderer_set_scene": (cint, [vp, vp, sz]),
        "pm_renderer_set_scene_device": (cint, [vp, vp, sz]),
        "pm_renderer_render": (cint, [vp]),
        "pm_renderer_set_frame_events": (cint, [vp, cint]),
        "pm_renderer_sync": (cint, [vp, ctypes.POINTER(FrameStats)]),
        "pm_renderer_read_rgba8": (cint, [vp, vp, sz]),
        "pm_renderer_render_host": (cint, [vp, vp, sz, vp, sz, ctypes.POINTER(FrameStats)]),
        "pm_renderer_framebuffer": (cint, [vp, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(u32)]),
        "pm_renderer_stream": (cint, [vp, ctypes.POINTER(vp)]),
        "pm_renderer_read_rgba32f": (cint, [vp, vp, sz]),
        "pm_renderer_read_tile_items": (cint, [vp, vp, vp, sz, ctypes.POINTER(sz), vp]),
        "pm_host_alloc": (cint, [ctypes.POINTER(vp), sz]),
        "pm_host_free": (None, [vp]),
        "pm_renderer_set_scene_paths": (cint, [vp, ctypes.POINTER(PathSetC), dbl, dbl]),
        "pm_renderer_read_scene": (cint, [vp, vp, sz, ctypes.POINTER(sz)]),
        "pm_context_new": (cint, [ctypes.POINTER(vp), vp, u32, u32]),
        "pm_context_free": (None, [vp]),
        "pm_context_save": (cint, [vp]),
        "pm_context_restore": (cint, [vp]),
        "pm_context_transform": (cint, [vp, vp]),
        "pm_context_clear": (cint, [vp, u32]),
        "pm_context_fill": (cint, [vp, vp, sz, u32]),
        "pm_context_fill_even_odd": (cint, [vp, vp, sz, u32]),
        "pm_context_stroke": (cint, [vp, vp, sz, u32, dbl]),
        "pm_context_item_count": (u32, [vp]),
        "pm_context_path_set": (cint, [vp, ctypes.POINTER(PathSetC)]),
        "pm_context_finish": (cint, [vp, dbl]),
        "pm_group_create": (cint, [ctypes.POINTER(vp), vp, u32, u32]),
        "pm_group_destroy": (None, [vp]),
        "pm_group_size": (u32, [vp]),
        "pm_group_member": (cint, [vp, u32, ctypes.POINTER(vp)]),
        "pm_group_resize": (cint, [vp, u32, u32]),
        "pm_group_set_scene": (cint, [vp, vp, sz]),
        "pm_group_strip_bounds": (cint, [vp, vp, sz]),
        "pm_group_set_frame_events": (cint, [vp, cint]),
        "pm_group_render": (cint, [vp]),
        "pm_group_sync": (cint, [vp, vp, sz, ctypes.POINTER(flt)]),
        "pm_group_read_rgba8": (cint, [vp, vp, sz]),
        "pm_group_gather_device": (cint, [vp, u32, ctypes.POINTER(vp), ctypes.POINTER(sz)]),
        "pm_group_nccl_version": (cint, []),
    }
    for name in EXPORTS:
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype, fn.argtypes = sig[name]
    _LIB = lib
    return lib


def _check(status, where):
    if status != PM_OK:
        raise PietMetalError(status, where)


def _ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


def version():
    return _lib().pm_version().decode()


# ------------------------------------------------------------------------------------------------
# feed
# ------------------------------------------------------------------------------------------------
def build_scene(kind, width, height, count=0, seed=0, scale=0.0, rect=(0.0, 0.0, 0.0, 0.0), rgba=0, options=0):
    """pm_scene_build: returns the encoded scene as a uint8 numpy array."""
    lib = _lib()
    d = SceneDesc(kind=kind, width=width, height=height, count=count, seed=seed, scale=scale, rgba=rgba, options=options)
    for i in range(4):
        d.rect[i] = float(rect[i])
    need = lib.pm_scene_build(ctypes.byref(d), None, 0)
    if need < 0:
        raise PietMetalError(int(need), "pm_scene_build")
    buf = np.zeros(int(need), np.uint8)
    got = lib.pm_scene_build(ctypes.byref(d), _ptr(buf), buf.size)
    if got != need:
        raise PietMetalError(int(got) if got < 0 else PM_ERR_BUFFER_TOO_SMALL, "pm_scene_build")
    return buf


def init_test_scene(buf_size=16 * 1024 * 1024):
    """The reference's C entry point (include/piet_metal.h:3) into a caller-owned buffer."""
    buf = np.zeros(buf_size, np.uint8)
    _lib().init_test_scene(_ptr(buf), buf_size)
    return buf


def scene_len(buf):
    """Bytes actually used by an encoded scene (highest ref in it)."""
    n = int(buf[:4].view(np.uint32)[0])
    items_ix = int(buf[4:8].view(np.uint32)[0])
    end = items_ix + 32 * n
    items = buf[items_ix:items_ix + 32 * n].view(np.uint32).reshape(n, 8)
    for tag, npts, pix in ((3, 3, 4), (4, 3, 4)):
        sel = items[:, 0] == tag
        if sel.any():
            end = max(end, int((items[sel, pix].astype(np.int64) + 8 * items[sel, npts].astype(np.int64)).max()))
    return end


def validate_scene(buf):
    return _lib().pm_scene_validate(_ptr(buf), buf.size)


def parse_color(s):
    return _lib().pm_parse_color(s.encode())


def flatten_svg_path(d, scale=1.0, tolerance=0.1):
    """flatten_path(BezPath::from_svg(d) scaled): list of (n_i, 2) float64 arrays, one per subpath."""
    lib = _lib()
    need = ctypes.c_size_t(0)
    cap_pts, cap_sub = 1 << 16, 1 << 12
    while True:
        xy = np.zeros((cap_pts, 2), np.float64)
        counts = np.zeros(cap_sub, np.uint32)
        n = lib.pm_flatten_svg_path(d.encode(), scale, tolerance, _ptr(xy), cap_pts, _ptr(counts), cap_sub, ctypes.byref(need))
        if n == PM_ERR_BUFFER_TOO_SMALL:
            cap_pts, cap_sub = max(cap_pts * 2, need.value), cap_sub * 2
            continue
        if n < 0:
            raise PietMetalError(int(n), "pm_flatten_svg_path")
        out, k = [], 0
        for i in range(int(n)):
            out.append(xy[k:k + counts[i]].copy())
            k += int(counts[i])
        return out


class Encoder:
    """Mirror of the reference's `Encoder` (src/lib.rs:79-254) over a caller-sized buffer."""

    def __init__(self, capacity):
        self.buf = np.zeros(capacity, np.uint8)
        self._h = ctypes.c_void_p()
        _check(_lib().pm_encoder_new(ctypes.byref(self._h), _ptr(self.buf), capacity), "pm_encoder_new")

    def begin_group(self, n_items):
        _check(_lib().pm_encoder_begin_group(self._h, n_items), "begin_group")

    def end_group(self):
        _check(_lib().pm_encoder_end_group(self._h), "end_group")

    def circle(self, cx, cy, r):
        _check(_lib().pm_encoder_circle(self._h, cx, cy, r), "circle")

    def stroke_line(self, p0, p1, width, rgba):
        _check(_lib().pm_encoder_stroke_line(self._h, p0[0], p0[1], p1[0], p1[1], width, rgba), "stroke_line")

    def fill(self, points, rgba, flags=None):
        pts = np.ascontiguousarray(points, np.float64)
        if flags is None:
            _check(_lib().pm_encoder_fill(self._h, _ptr(pts), pts.shape[0], rgba), "fill")
        else:
            _check(_lib().pm_encoder_fill_rule(self._h, _ptr(pts), pts.shape[0], rgba, flags), "fill_rule")

    def fill_subpaths(self, subpaths, rgba, flags=0):
        """One Fill item for a path of several closed subpaths (holes are cut out, not painted over)."""
        pts = np.ascontiguousarray(np.concatenate([np.asarray(sp, np.float64).reshape(-1, 2) for sp in subpaths]), np.float64)
        counts = np.ascontiguousarray([len(sp) for sp in subpaths], np.uint32)
        _check(_lib().pm_encoder_fill_subpaths(self._h, _ptr(pts), _ptr(counts), counts.size, rgba, flags), "fill_subpaths")

    def polyline(self, points, rgba, width):
        pts = np.ascontiguousarray(points, np.float64)
        _check(_lib().pm_encoder_polyline(self._h, _ptr(pts), pts.shape[0], rgba, width), "polyline")

    def bytes(self):
        """The encoded scene (a copy of the used prefix of the buffer)."""
        return self.buf[:_lib().pm_encoder_bytes(self._h)].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().pm_encoder_free(self._h)
            self._h = None


# ------------------------------------------------------------------------------------------------
# renderer
# ------------------------------------------------------------------------------------------------
class PietRenderer:
    """Mirror of the reference's PietRenderer (TestApp/PietRenderer.m) on one B200."""

    def __init__(self, device=0, flags=0, scratch_bytes=0):
        self._h = ctypes.c_void_p()
        cfg = Config(device=device, flags=flags, scratch_bytes=scratch_bytes)
        _check(_lib().pm_renderer_create(ctypes.byref(self._h), ctypes.byref(cfg)), "pm_renderer_create")
        self.width = self.height = 0
        self.tile_y0 = self.tile_y1 = 0

    def close(self):
        if getattr(self, "_h", None):
            _lib().pm_renderer_destroy(self._h)
            self._h = None

    __del__ = close

    # mtkView:drawableSizeWillChange: (PietRenderer.m:105)
    def drawable_size_will_change(self, width, height):
        _check(_lib().pm_renderer_resize(self._h, width, height), "pm_renderer_resize")
        self.width, self.height = width, height
        self.tile_y0, self.tile_y1 = 0, (height + 15) // 16

    resize = drawable_size_will_change

    def set_strip(self, tile_y0, tile_y1):
        _check(_lib().pm_renderer_set_strip(self._h, tile_y0, tile_y1), "pm_renderer_set_strip")
        self.tile_y0, self.tile_y1 = tile_y0, tile_y1

    # initScene (PietRenderer.m:203): hand the encoded scene to the renderer
    def init_scene(self, scene):
        scene = np.ascontiguousarray(scene, np.uint8)
        _check(_lib().pm_renderer_set_scene(self._h, _ptr(scene), scene.size), "pm_renderer_set_scene")

    set_scene = init_scene

    def set_scene_paths(self, path_set, scale=1.0, tolerance=0.1):
        """Flatten and encode a PathSet on the device (src/flatten.rs + Encoder::fill / polyline, which the reference runs on the CPU)."""
        c, keep = path_set.arrays()
        _check(_lib().pm_renderer_set_scene_paths(self._h, ctypes.byref(c), scale, tolerance), "pm_renderer_set_scene_paths")
        del keep

    def read_scene(self):
        n = ctypes.c_size_t(0)
        _lib().pm_renderer_read_scene(self._h, None, 0, ctypes.byref(n))
        buf = np.empty(n.value, np.uint8)
        _check(_lib().pm_renderer_read_scene(self._h, _ptr(buf), buf.size, ctypes.byref(n)), "pm_renderer_read_scene")
        return buf

    def set_scene_device(self, dev_ptr, nbytes):
        _check(_lib().pm_renderer_set_scene_device(self._h, ctypes.c_void_p(dev_ptr), nbytes), "pm_renderer_set_scene_device")

    # drawInMTKView: (PietRenderer.m:59): enqueue one frame
    def draw(self):
        _check(_lib().pm_renderer_render(self._h), "pm_renderer_render")

    render = draw

    def set_frame_events(self, enabled):
        """0 / False: no events, frames overlap; 1 / True: every kernel timed (serial); 2: the frame timed, its kernels overlap."""
        _check(_lib().pm_renderer_set_frame_events(self._h, int(enabled)), "pm_renderer_set_frame_events")

    def sync(self):
        st = FrameStats()
        _check(_lib().pm_renderer_sync(self._h, ctypes.byref(st)), "pm_renderer_sync")
        return st

    @property
    def strip_rows(self):
        return min(self.tile_y1 * 16, self.height) - self.tile_y0 * 16

    def read_rgba8(self, out=None):
        if out is None:
            out = np.empty((self.strip_rows, self.width, 4), np.uint8)
        _check(_lib().pm_renderer_read_rgba8(self._h, _ptr(out), out.strides[0]), "pm_renderer_read_rgba8")
        return out

    def read_rgba32f(self):
        out = np.empty((self.strip_rows, self.width, 4), np.float32)
        _check(_lib().pm_renderer_read_rgba32f(self._h, _ptr(out), out.strides[0]), "pm_renderer_read_rgba32f")
        return out

    def read_tile_items(self):
        """(offsets[n_tiles+1], items[structured], solid_colors[n_tiles]) of the strip, row-major."""
        n_tiles = (self.tile_y1 - self.tile_y0) * ((self.width + 15) // 16)
        offsets = np.zeros(n_tiles + 1, np.uint32)
        solid = np.zeros(n_tiles, np.uint32)
        cap = max(1024, 4 * n_tiles)
        while True:
            items = np.zeros(cap, TILE_ITEM_DTYPE)
            n = ctypes.c_size_t(0)
            st = _lib().pm_renderer_read_tile_items(self._h, _ptr(offsets), _ptr(items), cap, ctypes.byref(n), _ptr(solid))
            if st == PM_ERR_BUFFER_TOO_SMALL:
                cap = n.value
                continue
            _check(st, "pm_renderer_read_tile_items")
            return offsets, items[:n.value], solid

    def render_host(self, scene, out=None):
        """Host bytes in, host pixels out: upload + one frame + read-back (the e2e call)."""
        scene = np.ascontiguousarray(scene, np.uint8)
        if out is None:
            out = np.empty((self.strip_rows, self.width, 4), np.uint8)
        st = FrameStats()
        _check(_lib().pm_renderer_render_host(self._h, _ptr(scene), scene.size, _ptr(out), out.strides[0], ctypes.byref(st)),
               "pm_renderer_render_host")
        return out, st

    def framebuffer(self):
        """(device pointer, pitch in bytes, rows) of the strip's RGBA8 framebuffer."""
        p, pitch, rows = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_uint32()
        _check(_lib().pm_renderer_framebuffer(self._h, ctypes.byref(p), ctypes.byref(pitch), ctypes.byref(rows)), "pm_renderer_framebuffer")
        return p.value, pitch.value, rows.value

    def stream(self):
        s = ctypes.c_void_p()
        _check(_lib().pm_renderer_stream(self._h, ctypes.byref(s)), "pm_renderer_stream")
        return s.value or 0


def row_costs(scene, width, height):
    """pm_scene_row_costs: relative cost of every tile row of the frame (float32 array)."""
    scene = np.ascontiguousarray(scene, np.uint8)
    n = (height + 15) // 16
    cost = np.zeros(n, np.float32)
    _check(_lib().pm_scene_row_costs(_ptr(scene), scene.size, width, height, _ptr(cost), n), "pm_scene_row_costs")
    return cost


def balanced_strip_bounds(cost, world_size):
    """pm_balance_strips: contiguous non-empty strips of tile rows with the smallest possible maximum cost."""
    cost = np.ascontiguousarray(cost, np.float32)
    bounds = np.zeros(world_size + 1, np.uint32)
    _check(_lib().pm_balance_strips(_ptr(cost), cost.size, world_size, _ptr(bounds)), "pm_balance_strips")
    return [int(b) for b in bounds]


def write_image(path, rgba8):
    """pm_write_png / pm_write_ppm (by extension) of an (H, W, 4) uint8 array."""
    img = np.ascontiguousarray(rgba8, np.uint8)
    fn = _lib().pm_write_ppm if str(path).lower().endswith(".ppm") else _lib().pm_write_png
    _check(fn(str(path).encode(), _ptr(img), img.shape[1], img.shape[0], img.strides[0]), "pm_write_image")


def strip_bounds(n_tile_rows, world_size):
    """Contiguous row-strip shard of the frame's tile rows: rank g renders [b[g], b[g+1])."""
    return [(n_tile_rows * g) // world_size for g in range(world_size + 1)]


class PathSet:
    """Path control points for pm_renderer_set_scene_paths (flattening + encoding on the device): one item per subpath."""

    def __init__(self):
        self.first, self.start, self.verb, self.ctrl = [0], [], [], []
        self.tag, self.rgba, self.width, self.flags = [], [], [], []

    def begin(self, x, y, tag, rgba, width=0.0, flags=0):
        """MoveTo: a new subpath = a new item (tag 3 = PietFill, 4 = PietStrokePolyLine)."""
        self.start.append((float(x), float(y)))
        self.tag.append(tag); self.rgba.append(rgba); self.width.append(width); self.flags.append(flags)
        self.first.append(self.first[-1])

    def line_to(self, x, y):
        self.verb.append(0); self.ctrl.append((0.0, 0.0, 0.0, 0.0, float(x), float(y))); self.first[-1] += 1

    def curve_to(self, c1x, c1y, c2x, c2y, x, y):
        self.verb.append(1); self.ctrl.append((float(c1x), float(c1y), float(c2x), float(c2y), float(x), float(y))); self.first[-1] += 1

    def arrays(self):
        a = {"first": np.ascontiguousarray(self.first, np.uint32), "start": np.ascontiguousarray(self.start, np.float64).reshape(-1, 2),
             "verb": np.ascontiguousarray(self.verb, np.uint8), "ctrl": np.ascontiguousarray(self.ctrl, np.float64).reshape(-1, 6),
             "tag": np.ascontiguousarray(self.tag, np.uint32), "rgba": np.ascontiguousarray(self.rgba, np.uint32),
             "width": np.ascontiguousarray(self.width, np.float32), "flags": np.ascontiguousarray(self.flags, np.uint32)}
        c = PathSetC(n_subpaths=len(self.tag), n_segments=len(self.verb), first_segment=a["first"].ctypes.data, start=a["start"].ctypes.data,
                     verb=a["verb"].ctypes.data if len(self.verb) else None, ctrl=a["ctrl"].ctypes.data if len(self.verb) else None,
                     tag=a["tag"].ctypes.data, rgba=a["rgba"].ctypes.data, width=a["width"].ctypes.data, flags=a["flags"].ctypes.data)
        return c, a  # (keep `a` alive while `c` is in use)


PATH_EL_DTYPE = np.dtype([("verb", np.uint32), ("pad", np.uint32), ("x", np.float64, 6)])


class BezPath:
    """kurbo-style path builder for RenderContext."""

    def __init__(self):
        self.els = []

    def move_to(self, x, y): self.els.append((0, (x, y, 0, 0, 0, 0))); return self
    def line_to(self, x, y): self.els.append((1, (x, y, 0, 0, 0, 0))); return self
    def quad_to(self, cx, cy, x, y): self.els.append((2, (cx, cy, x, y, 0, 0))); return self
    def curve_to(self, c1x, c1y, c2x, c2y, x, y): self.els.append((3, (c1x, c1y, c2x, c2y, x, y))); return self
    def close_path(self): self.els.append((4, (0, 0, 0, 0, 0, 0))); return self

    def array(self):
        a = np.zeros(len(self.els), PATH_EL_DTYPE)
        for i, (v, x) in enumerate(self.els):
            a[i]["verb"] = v
            a[i]["x"] = x
        return a


class RenderContext:
    """The piet RenderContext calls over a PietRenderer: clear, transform, save / restore, fill, fill_even_odd, stroke
    (solid colours 0xRRGGBBAA), finish -- which installs the drawing as the renderer's scene, flattened and encoded on the
    device.  Create the renderer with FLAG_FILL_RULES for fill_even_odd to be honoured."""

    def __init__(self, renderer, width=None, height=None):
        self.renderer = renderer
        self._h = ctypes.c_void_p()
        _check(_lib().pm_context_new(ctypes.byref(self._h), renderer._h if renderer is not None else None,
                                     width or renderer.width, height or renderer.height), "pm_context_new")

    def close(self):
        if getattr(self, "_h", None):
            _lib().pm_context_free(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def solid_brush(rgba):
        return rgba

    def clear(self, rgba): _check(_lib().pm_context_clear(self._h, rgba), "clear")
    def save(self): _check(_lib().pm_context_save(self._h), "save")
    def restore(self): _check(_lib().pm_context_restore(self._h), "restore")

    def transform(self, affine):
        m = np.ascontiguousarray(affine, np.float64)
        _check(_lib().pm_context_transform(self._h, _ptr(m)), "transform")

    def fill(self, path, brush):
        a = path.array()
        _check(_lib().pm_context_fill(self._h, _ptr(a), a.size, brush), "fill")

    def fill_even_odd(self, path, brush):
        a = path.array()
        _check(_lib().pm_context_fill_even_odd(self._h, _ptr(a), a.size, brush), "fill_even_odd")

    def stroke(self, path, brush, width):
        a = path.array()
        _check(_lib().pm_context_stroke(self._h, _ptr(a), a.size, brush, width), "stroke")

    def item_count(self):
        return _lib().pm_context_item_count(self._h)

    def path_set(self):
        """The recorded drawing as numpy arrays (copies): first, start, verb, ctrl, tag, rgba, width, flags."""
        c = PathSetC()
        _check(_lib().pm_context_path_set(self._h, ctypes.byref(c)), "path_set")
        ns, ng = c.n_subpaths, c.n_segments

        def arr(ptr, dtype, n):
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy() if n else np.zeros(0, dtype)
        return {"first": arr(c.first_segment, np.uint32, ns + 1), "start": arr(c.start, np.float64, 2 * ns).reshape(-1, 2),
                "verb": arr(c.verb, np.uint8, ng), "ctrl": arr(c.ctrl, np.float64, 6 * ng).reshape(-1, 6), "tag": arr(c.tag, np.uint32, ns),
                "rgba": arr(c.rgba, np.uint32, ns), "width": arr(c.width, np.float32, ns), "flags": arr(c.flags, np.uint32, ns)}

    def finish(self, tolerance=0.1):
        _check(_lib().pm_context_finish(self._h, tolerance), "finish")


class PietRendererGroup:
    """N GPUs of one box behind one handle (pm_group_*): the scene is uploaded once and broadcast with NCCL inside
    the library; every GPU renders one contiguous, cost-balanced strip of tile rows."""

    def __init__(self, devices, flags=0):
        devs = np.ascontiguousarray(list(devices), np.int32)
        self._h = ctypes.c_void_p()
        _check(_lib().pm_group_create(ctypes.byref(self._h), _ptr(devs), devs.size, flags), "pm_group_create")
        self.n = devs.size
        self.width = self.height = 0

    def close(self):
        if getattr(self, "_h", None):
            _lib().pm_group_destroy(self._h)
            self._h = None

    __del__ = close

    def resize(self, width, height):
        _check(_lib().pm_group_resize(self._h, width, height), "pm_group_resize")
        self.width, self.height = width, height

    def set_scene(self, scene):
        scene = np.ascontiguousarray(scene, np.uint8)
        _check(_lib().pm_group_set_scene(self._h, _ptr(scene), scene.size), "pm_group_set_scene")

    def strip_bounds(self):
        b = np.zeros(self.n + 1, np.uint32)
        _check(_lib().pm_group_strip_bounds(self._h, _ptr(b), b.size), "pm_group_strip_bounds")
        return [int(x) for x in b]

    def set_frame_events(self, mode):
        _check(_lib().pm_group_set_frame_events(self._h, int(mode)), "pm_group_set_frame_events")

    def render(self):
        _check(_lib().pm_group_render(self._h), "pm_group_render")

    def sync(self):
        """(list of FrameStats per member, slowest member's last frame in ms)"""
        stats = (FrameStats * self.n)()
        worst = ctypes.c_float(0)
        _check(_lib().pm_group_sync(self._h, ctypes.cast(stats, ctypes.c_void_p), self.n, ctypes.byref(worst)), "pm_group_sync")
        return list(stats), worst.value

    def read_rgba8(self):
        out = np.empty((self.height, self.width, 4), np.uint8)
        _check(_lib().pm_group_read_rgba8(self._h, _ptr(out), out.strides[0]), "pm_group_read_rgba8")
        return out

    def gather_device(self, root=0):
        """(device pointer, pitch in bytes) of the whole frame gathered on member `root`'s GPU."""
        ptr, pitch = ctypes.c_void_p(), ctypes.c_size_t(0)
        _check(_lib().pm_group_gather_device(self._h, root, ctypes.byref(ptr), ctypes.byref(pitch)), "pm_group_gather_device")
        return ptr.value, pitch.value


def nccl_version():
    return _lib().pm_group_nccl_version()
