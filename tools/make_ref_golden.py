"""Writes tests/golden/ref_vectors.json: digests of what the REFERENCE ITSELF produces for the pinned
cases of tests/test_ref_pin.py -- its unmodified PietRender.metal run through oracle/_ref (built by
`make -C oracle ref` from /root/reference; only possible in the dev container).  The oracle has to
reproduce these digests (test_oracle_matches_reference_golden), so the pin travels with the repository
even where neither /root/reference nor the built library exists.

    python tools/make_ref_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_api  # noqa: E402
import test_ref_pin as pin  # noqa: E402


def main():
    pm = ge.load_package()
    if not oracle_api.have_ref():
        raise SystemExit("oracle/_ref/libpm_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {"generator": "tools/make_ref_golden.py", "source": "linebender/piet-metal @ 71afb99, TestApp/PietRender.metal via oracle/metal_shim",
           "cases": {}}
    for name, scene, w, h in pin.pin_cases(pm):
        ref = oracle_api.ref_render(scene, w, h, want_cmds=True)
        trusted = oracle_api.ref_trusted_tiles(ref["n_cmds"])
        tiles = pin.stream_tiles(trusted)
        streams = {int(t): (oracle_api.canonical_cmds(ref["cmds"][t][:int(ref["n_cmds"][t]) * 24]), int(ref["solid"][t])) for t in tiles}
        d = pin.digest_case(oracle_api, ref, scene, w, h, trusted, streams)
        d["scene_sha256"] = hashlib.sha256(scene.tobytes()).hexdigest()
        d["width"], d["height"] = w, h
        d["untrusted_tiles"] = [int(t) for t in np.flatnonzero(~trusted)]
        d["stream_tiles"] = [int(t) for t in tiles]
        d["max_cmds_per_tile"] = int(ref["n_cmds"].max())
        out["cases"][name] = d
        print("%-28s %5dx%-5d untrusted tiles %d, max cmds/tile %d" % (name, w, h, len(d["untrusted_tiles"]), d["max_cmds_per_tile"]))
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
