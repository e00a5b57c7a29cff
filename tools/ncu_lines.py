#!/usr/bin/env python3
"""Per-source-line attribution of an ncu source-page CSV (SASS view) using nvdisasm line info.

    tools/ncu_lines.py <source.csv> <cubin> <kernel-substring> [top]

Joins the two instruction streams by order (opcodes are checked) and sums executed warp
instructions and stall samples per file:line."""
import csv, re, subprocess, sys, collections, os

def _section(rows, kern):
    """rows of the source-page CSV that belong to the kernel whose name contains `kern` (a file may hold several)."""
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        return rows
    pick = None
    for k, i in enumerate(starts):
        name = rows[i][1] if len(rows[i]) > 1 else ""
        short = kern.split("E")[0].split("IL")[0]
        if short in name.replace("::", ""):
            pick = k
            break
    if pick is None:
        pick = 0
    end = starts[pick + 1] if pick + 1 < len(starts) else len(rows)
    return rows[starts[pick]:end]

src_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = _section(list(csv.reader(open(src_csv))), kern)
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
prof = [r for r in rows[hdr_i + 1:] if len(r) > col["Instructions Executed"]]

dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the section of the kernel
insts = []  # (opcode text, file, line)
active = False
cur = ("?", 0)
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        active = kern in m.group(1)
        continue
    if re.match(r"\s*\.section", ln):
        active = False
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        insts.append((m.group(2).strip(), cur[0], cur[1]))
print("profile instructions: %d, disassembly: %d" % (len(prof), len(insts)))
n = min(len(prof), len(insts))
mism = 0
agg = collections.defaultdict(lambda: [0, 0, 0])
total_i = total_s = 0
for k in range(n):
    p = prof[k]
    op_p = p[col["Source"]].split()[0:2]
    op_d = insts[k][0].split()[0:2]
    if op_p and op_d and op_p[0].lstrip("@!P0123456789UT ") != op_d[0].lstrip("@!P0123456789UT "):
        if (op_p[-1] if op_p[0].startswith("@") else op_p[0]).split(".")[0] != (op_d[-1] if op_d[0].startswith("@") else op_d[0]).split(".")[0]:
            mism += 1
    ie = int(p[col["Instructions Executed"]] or 0)
    ss = int(p[col["# Samples"]] or 0)
    key = (insts[k][1], insts[k][2])
    agg[key][0] += ie; agg[key][1] += ss; agg[key][2] += 1
    total_i += ie; total_s += ss
print("opcode mismatches: %d; total warp instructions %d, samples %d" % (mism, total_i, total_s))
items = sorted(agg.items(), key=lambda kv: -kv[1][0])
print("%-28s %12s %6s %8s %6s %5s" % ("file:line", "warp inst", "%", "samples", "%", "SASS"))
for (f, l), (ie, ss, cnt) in items[:top]:
    print("%-28s %12d %6.2f %8d %6.2f %5d" % ("%s:%d" % (f, l), ie, 100.0 * ie / max(1, total_i), ss, 100.0 * ss / max(1, total_s), cnt))
if os.environ.get("BY_RANGE"):
    # ranges "name:file:lo-hi,..." -> sums
    for spec in os.environ["BY_RANGE"].split(","):
        name, f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        ie = sum(v[0] for (ff, l), v in agg.items() if ff == f and lo <= l <= hi)
        ss = sum(v[1] for (ff, l), v in agg.items() if ff == f and lo <= l <= hi)
        print("%-24s inst %6.2f%%  samples %6.2f%%" % (name, 100.0 * ie / total_i, 100.0 * ss / total_s))
