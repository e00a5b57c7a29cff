#!/usr/bin/env python3
"""Writes the committed golden fixtures under tests/golden/ from the oracle.

The reference ships no golden vectors and cannot run here, so these pin the
oracle against regressions, not against reference output; the hand-derived known answers in
tests/test_oracle_kat.py are what anchors its semantics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_api  # noqa: E402

pm = ge.load_package()
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
scene = pm.build_scene(pm.SCENE_TIGER, 128, 128)
np.save(os.path.join(out, "tiger_128_rgba8.npy"), oracle_api.render(scene, 128, 128)["rgba8"])
scene = pm.build_scene(pm.SCENE_CARDIOID, 512, 384, scale=0.25)
res = oracle_api.render(scene, 512, 384, items=True)
np.savez_compressed(os.path.join(out, "cardioid_512x384.npz"), rgba8=res["rgba8"], offsets=res["offsets"],
                    items=res["items"], solid=res["solid"])
print("golden fixtures written to", out)
