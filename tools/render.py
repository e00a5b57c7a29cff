#!/usr/bin/env python3
"""Render one of the built-in scenes to a PNG / PPM.

    tools/render.py tiger 2048 out.png            # through the CUDA library (needs a B200)
    tools/render.py tiger 512 out.png --oracle    # through the CPU oracle (test infrastructure; for eyeballing only)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge

pm = ge.load_package()
kinds = {"tiger": pm.SCENE_TIGER, "cardioid": pm.SCENE_CARDIOID, "path_test": pm.SCENE_PATH_TEST,
         "rand_bezier": pm.SCENE_RAND_BEZIER, "glyphs": pm.SCENE_GLYPHS}
name, size, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
scene = pm.build_scene(kinds[name], size, size)
if "--oracle" in sys.argv:
    import oracle_api
    img = oracle_api.render(scene, size, size)["rgba8"]
else:
    r = pm.PietRenderer(device=0)
    r.drawable_size_will_change(size, size)
    img, _ = r.render_host(scene)
    r.close()
pm.write_image(out, img)
print("wrote %s (%dx%d)" % (out, size, size))
