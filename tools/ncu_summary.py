#!/usr/bin/env python3
"""Key metrics of an ncu raw-page CSV (one kernel)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_active.avg.per_cycle_active", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
d = dict(zip(hdr, vals))
for k in want:
    if k in d: print("%-70s %s" % (k, d[k]))
for k in hdr:
    if "issue_stalled" in k and "per_issue_active" in k:
        v = float(d[k])
        if v > 0.1: print("  stall %-40s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v))
