#!/usr/bin/env python3
"""condense one kernel of an .ncu-rep into the handful of numbers the design notes quote:
   python tools/ncu_summary.py report.ncu-rep [more metric substrings]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct", "sm__inst_executed_pipe_fp64", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct", "lts__t_bytes.sum ", "l1tex__t_bytes.sum ", "sm__icc_requests.sum ", "sm__icc_request_hit_rate",
        "smsp__inst_executed_op_branch.sum", "local_op_ld.sum ", "local_op_st.sum ", "smsp__average_warps_issue_stalled",
        "shared_op_ld.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "smsp__inst_executed_op_shared"]
def main():
    rep = sys.argv[1]; extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[-1]
    for k, un, x in zip(h, u, v):
        kk = k + " "
        if any(t in kk for t in KEYS + extra):
            if "stalled" in k and "per_issue_active" not in k: continue
            if "stalled" in k and float(x or 0) < 0.3: continue
            print(f"{k:90s} {x} {un}")
main()
