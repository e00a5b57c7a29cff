#!/usr/bin/env python3
"""Per-function attribution (innermost inlined function by line) of an ncu source-page CSV.
    tools/ncu_funcs.py <source.csv> <cubin> <kernel-substring> <src-root>"""
import csv, re, subprocess, sys, os, collections, bisect

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
src_csv, cubin, kern, root = sys.argv[1:5]
rows = _section(list(csv.reader(open(src_csv))), kern)
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hi])}
prof = [r for r in rows[hi + 1:] if len(r) > col["Instructions Executed"]]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
insts, active, cur = [], False, ("?", 0)
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m: active = kern in m.group(1); continue
    if re.match(r"\s*\.section", ln): active = False
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln): insts.append(cur)
# function start lines per file
starts = {}
for f in set(i[0] for i in insts):
    path = os.path.join(root, f)
    if not os.path.exists(path): continue
    lst = []
    for n, line in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:PM_HD|__device__|__global__|static|inline).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
        if m and not line.startswith(" "): lst.append((n, m.group(1)))
        m2 = re.match(r"^\s+// \[section: (.*)\]", line)
        if m2: lst.append((n, "  §" + m2.group(1)))
    starts[f] = lst
def func(f, l):
    lst = starts.get(f)
    if not lst: return f
    k = bisect.bisect_right([x[0] for x in lst], l) - 1
    return "%s:%s" % (f.split(".")[0][-8:], lst[k][1]) if k >= 0 else f
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
ti = ts = 0
for k in range(min(len(prof), len(insts))):
    ie = int(prof[k][col["Instructions Executed"]] or 0); ss = int(prof[k][col["# Samples"]] or 0)
    a = agg[func(*insts[k])]; a[0] += ie; a[1] += ss; a[2] += 1; a[3] += 1 if ie >= 10000 else 0
    ti += ie; ts += ss
print("total warp inst %d, samples %d, static %d" % (ti, ts, len(insts)))
print("%-34s %7s %7s %6s %6s" % ("function", "inst%", "smpl%", "SASS", "hot"))
for name, (ie, ss, n, hot) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if ie * 1000 < ti and n < 30: continue
    print("%-34s %7.2f %7.2f %6d %6d" % (name, 100.0 * ie / ti, 100.0 * ss / ts, n, hot))
