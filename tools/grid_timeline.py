#!/usr/bin/env python3
"""Where a frame's time goes, from %globaltimer marks inside the kernels (PM_DEBUG_FINE=1 is set here; k_fine's per-warp marks need
a library built with -DPM_FINE_TIMELINE=1: make -C piet-metal_b200 variant NAME=tl VFLAGS=-DPM_FINE_TIMELINE=1).
    PM_LIB=.../variants/libpm_tl.so tools/grid_timeline.py <size> [N:g]"""
import ctypes, os, sys
os.environ["PM_DEBUG_FINE"] = "1"  # allocates the debug buffer (PM_DEBUG_SEG would also make sync() print and reset it)
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pm = ge.load_package()
size = int(sys.argv[1])
scene = pm.build_scene(pm.SCENE_TIGER, size, size)
r = pm.PietRenderer(device=0)
r.drawable_size_will_change(size, size)
if len(sys.argv) > 2:
    n, g = [int(x) for x in sys.argv[2].split(":")]
    b = pm.balanced_strip_bounds(pm.row_costs(scene, size, size), n)
    r.set_strip(b[g], b[g + 1])
r.init_scene(scene)
r.set_frame_events(0)
lib = pm._lib()
lib.pm_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
buf = np.zeros(1 << 20, np.uint64)
def read():
    assert lib.pm_debug_read(r._h, buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
    return buf.copy()
for _ in range(5): r.draw()
r.sync(); read()
r.draw(); r.sync()
d = read().astype(np.int64)
def ctas(base):
    a = d[base:base + (1 << 18)].reshape(-1, 2)
    return a[(a[:, 1] > 0) & (a[:, 0] > 0)]
seg, row = ctas(0), ctas(1 << 18)
t0 = seg[:, 0].min()
def show(name, a):
    if len(a) == 0: print(name, "no marks"); return
    s, e = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3
    life = e - s
    print("%-6s %5d CTAs: first start %6.1f us, last start %6.1f, last end %6.1f | CTA lifetime p10 %.1f p50 %.1f p90 %.1f max %.1f us" %
          (name, len(a), s.min(), s.max(), e.max(), *np.percentile(life, [10, 50, 90]), life.max()))
    ts = np.arange(0, e.max(), 2.0)
    print("       running CTAs every 2 us:", " ".join(str(int(((s <= t) & (e > t)).sum())) for t in ts))
show("k_seg", seg); show("k_row", row)
f = d[1 << 19:(1 << 19) + 8 * 148 * 4 * 8].reshape(-1, 8)
f = f[f[:, 7] != 0]
if len(f):
    st, first, last, ex = [(f[:, k] - t0) / 1e3 for k in range(4)]
    n_tiles_w = f[:, 5] & 0xffffffff
    has = n_tiles_w > 0
    print("k_fine %5d warps: entry %6.1f..%6.1f us, first tile starts %6.1f..%6.1f (median %.1f), last tile ends median %.1f max %.1f, exit max %.1f" %
          (len(f), st.min(), st.max(), first[has].min(), first[has].max(), np.median(first[has]), np.median(last[has]), last[has].max(), ex.max()))
    print("       tiles per warp median %d max %d; time in tiles per warp median %.1f us; longest tile %.1f us (p99 of per-warp longest %.1f)" %
          (np.median(n_tiles_w), n_tiles_w.max(), 0.0, (f[:, 4] >> 32).max() / 1e3, np.percentile(f[:, 4] >> 32, 99) / 1e3))
    ts = np.arange(st.min(), ex.max(), 4.0)
    print("       warps inside [first tile, last tile end] every 4 us:", " ".join(str(int((has & (first <= t) & (last > t)).sum())) for t in ts))
    # the end of the kernel in detail: warps still inside tiles every 2 us over the last 30 us, and the latest long tiles
    t_end = ex.max()
    ts = np.arange(t_end - 30.0, t_end, 2.0)
    print("       last 30 us, warps inside tiles every 2 us:", " ".join(str(int((has & (first <= t) & (last > t)).sum())) for t in ts))
    print("       CTAs' exits: p50 %.1f p90 %.1f p99 %.1f max %.1f us before the end" % tuple(t_end - np.percentile(ex, [50, 10, 1, 0])))
    dur = (f[:, 4] >> 32) / 1e3
    lstart = (f[:, 7] - t0) / 1e3
    lend = lstart + dur
    lstart2 = (f[:, 6] - t0) / 1e3
    order = np.argsort(-last)[:16]
    print("       the tiles that end last: (end, us before kernel end; duration us; tile row, col; tiles this warp drew)")
    for k in order:
        if not has[k]: continue
        e = int(f[k, 5] >> 32)
        print("         -%.1f us  %.1f us  (%d, %d)  %d tiles" % (t_end - last[k], last[k] - lstart2[k], e >> 16, e & 0xffff, n_tiles_w[k]))
