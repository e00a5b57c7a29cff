#!/usr/bin/env python3
"""Workload statistics of the fill kernel (CPU only, via the test harness of the device logic)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pm = ge.load_package()
h = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "libpm_host_harness.so"))
kind = {"tiger": pm.SCENE_TIGER, "rand_bezier": pm.SCENE_RAND_BEZIER, "glyphs": pm.SCENE_GLYPHS}[sys.argv[1]]
size = int(sys.argv[2])
scene = pm.build_scene(kind, size, size)
out = np.zeros(16, np.uint64); hi = np.zeros(32, np.uint64); hr = np.zeros(64, np.uint64)
y0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
y1 = int(sys.argv[4]) if len(sys.argv) > 4 else (size + 15) // 16
h.pmh_stats(scene.ctypes.data_as(ctypes.c_void_p), size, size, y0, y1, out.ctypes.data_as(ctypes.c_void_p),
            hi.ctypes.data_as(ctypes.c_void_p), hr.ctypes.data_as(ctypes.c_void_p))
names = ["tiles_with_records", "records", "items", "fill_pairs", "fill_near_px", "line_pairs", "line_px", "fill_edge_recs", "has_draw_tiles", "max_near_in_pair"]
for n, v in zip(names, out): print("%-20s %d" % (n, v))
t = float(out[8])
print("per drawn tile: records %.2f items %.2f fill pairs %.1f near %.1f line pairs %.1f line px %.1f" % (out[1]/t, out[2]/t, out[3]/t, out[4]/t, out[5]/t, out[6]/t))
print("items/tile hist:", [int(x) for x in hi])
print("records/tile hist:", [int(x) for x in hr])
