#!/usr/bin/env python3
"""Per-warp timeline of k_fine (debug build of the library with -DPM_FINE_TIMELINE=1, PM_DEBUG_FINE=1).
    PM_LIB=.../lib_timeline.so PM_DEBUG_FINE=1 tools/fine_timeline.py <size> [N:g]"""
import ctypes, os, sys
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
for _ in range(5):
    r.draw()
st = r.sync()
lib = pm._lib()
n_warps = 148 * 4 * 8
buf = np.zeros(n_warps * 24, np.uint64)
lib.pm_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
assert lib.pm_debug_read(r._h, buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
d = buf.reshape(n_warps, 24).astype(np.int64)
t0 = d[:, 0].min()
beg, lastc, end, tiles = (d[:, 0] - t0) / 1e3, (d[:, 1] - t0) / 1e3, (d[:, 2] - t0) / 1e3, d[:, 3]
longest, lpk = (d[:, 4] >> 32) / 1e3, d[:, 4] & 0xffffffff
print("frame %.1f us fine %.1f us; complex tiles %d; warps %d" % (st.ms_total * 1e3, st.ms_fine * 1e3, st.n_complex_tiles, n_warps))
print("warp begin: min %.1f max %.1f us" % (beg.min(), beg.max()))
print("last complex tile end: median %.1f p90 %.1f p99 %.1f max %.1f us" % tuple(np.percentile(lastc[tiles > 0], [50, 90, 99, 100])))
print("warp end: median %.1f p90 %.1f max %.1f us" % tuple(np.percentile(end, [50, 90, 100])))
print("tiles per warp (incl. skipped/no-draw): mean %.1f max %d; warps with tiles %d" % (tiles[tiles > 0].mean(), tiles.max(), (tiles > 0).sum()))
order = np.argsort(-longest)[:8]
for k in order:
    print("  longest tile of warp %5d: %.1f us  tile row %d col %d (ended at %.1f us, %d tiles)" % (k, longest[k], (lpk[k] >> 16), lpk[k] & 0xffff, lastc[k], tiles[k]))
print("sum of per-warp busy-with-complex estimates: longest-tile median %.2f us" % np.median(longest[tiles > 0]))
names = ["set-up", "item select", "phase A stroke (inline chunk)", "phase A overflow chunks", "resolve+blend", "encode+store", "no-draw/skipped", "phase A fill (inline chunk)"]
acc = d[:, 8:24].sum(axis=0) / 1e3
tot = acc.sum()
print("time by phase, summed over warps (us, %% of %.0f):" % tot)
for h in (0, 1):
    for k, nme in enumerate(names):
        v = acc[8 * h + k]
        if v > 0: print("  %-6s %-32s %9.0f  %5.1f%%" % ("heavy" if h else "light", nme, v, 100 * v / tot))
