#!/usr/bin/env python3
"""Top SASS instructions of an ncu source-page CSV by a stall column, with nvdisasm line info.
    tools/ncu_stalls.py <source.csv> <cubin> <kernel-substring> <column> [top]"""
import csv, re, subprocess, sys, os

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
src_csv, cubin, kern, column = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
rows = _section(list(csv.reader(open(src_csv))), kern)
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hdr_i])}
prof = [r for r in rows[hdr_i + 1:] if len(r) > col["Instructions Executed"]]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
insts, active, cur = [], False, ("?", 0)
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m: active = kern in m.group(1); continue
    if re.match(r"\s*\.section", ln): active = False
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: insts.append((m.group(2).strip(), cur[0], cur[1]))
tot = sum(int(p[col[column]] or 0) for p in prof)
order = sorted(range(len(prof)), key=lambda k: -int(prof[k][col[column]] or 0))
print("total %s = %d" % (column, tot))
for k in order[:top]:
    p = prof[k]
    print("%5d %6d %5.1f%%  %-22s %s" % (k, int(p[col[column]] or 0), 100.0 * int(p[col[column]] or 0) / max(1, tot),
                                        "%s:%d" % (insts[k][1], insts[k][2]), p[col["Source"]].strip()[:90]))
