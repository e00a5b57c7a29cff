#!/usr/bin/env python3
"""Derive the tiger input fixture from the reference's SVG test asset.

The reference embeds Ghostscript_Tiger.svg with include_bytes! (src/lib.rs:288) and walks the
<path> children of the first <g> (src/lib.rs:290-327), reading only the attributes d, fill, stroke
and stroke-width of each element.  This script extracts exactly those four attributes, in document
order, into a line-oriented "path list" that the C++ feed embeds:

    <fill|-> <stroke|-> <stroke-width|-> <path data>

Run in the authoring container only (the reference tree is not present on the GPU box):

    python tools/make_tiger_fixture.py /root/reference/Ghostscript_Tiger.svg \
        piet-metal_b200/assets/tiger.pathlist
"""
import re
import sys


def main(src, dst):
    text = open(src, encoding="utf-8").read()
    vb = re.search(r'viewBox="([^"]*)"', text).group(1)
    out = ["# derived from Ghostscript_Tiger.svg by tools/make_tiger_fixture.py", "viewbox " + vb]
    for attrs in re.findall(r"<path\s+([^>]*?)/?>", text, flags=re.S):
        kv = dict(re.findall(r'([\w-]+)="([^"]*)"', attrs))
        d = " ".join(kv["d"].split())
        out.append("path %s %s %s %s" % (kv.get("fill", "-"), kv.get("stroke", "-"),
                                          kv.get("stroke-width", "-"), d))
    open(dst, "w", encoding="utf-8").write("\n".join(out) + "\n")
    print("%d paths -> %s" % (len(out) - 2, dst))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
