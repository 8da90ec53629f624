#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion
needs.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-row-index] > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.sum.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__sass_inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shfl.sum",
    "sm__sass_thread_inst_executed_op_integer_pred_on.sum", "smsp__sass_thread_inst_executed_op_imad_pred_on.sum",
]
for k in KEYS:
    hit = [h for h in hdr if h == k] or [h for h in hdr if h.startswith(k)]
    for h in hit[:1]:
        print("%-75s %-12s %s" % (h, m[h][0], m[h][1]))
print("\n-- warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active, > 0.05) --")
st = [(float(m[h][1].replace(",", "")), h) for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h and m[h][1]]
for v, h in sorted(st, reverse=True):
    if v > 0.05:
        print("%-90s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
print("\n-- instruction mix (sm__sass_inst_executed / smsp__inst_executed_pipe_*) --")
for h in hdr:
    if h.startswith("smsp__inst_executed_pipe_") and h.endswith(".sum") and m[h][1] not in ("0", ""):
        print("%-75s %s" % (h, m[h][1]))
