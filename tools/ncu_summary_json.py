#!/usr/bin/env python3
"""Key metrics and stall reasons of every kernel in an ncu raw-page CSV, as JSON (what profiles/*_ncu_summary.json hold).
    tools/ncu_summary_json.py <raw.csv> <out.json>"""
import csv, json, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__icc_request_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__warps_active.avg.per_cycle_active", "sm__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    e = {"Kernel Name": d["Kernel Name"]}
    for k in WANT:
        if k in d: e[k] = ("%s %s" % (d[k], u[k])).strip()
    st = {}
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            try: v = float(d[k])
            except ValueError: continue
            if v > 0.1: st[k.split("issue_stalled_")[1].split("_per_issue")[0]] = round(v, 2)
    e["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
    res.append(e)
json.dump(res, open(sys.argv[2], "w"), indent=1)
for e in res:
    print("%-44s %12s %22s issue %s" % (e["Kernel Name"][:44], e.get("gpu__time_duration.sum"), e.get("smsp__inst_executed.sum"), e.get("smsp__issue_active.avg.pct_of_peak_sustained_active")))
