#!/usr/bin/env python
"""Print the handful of ncu metrics that matter for an HBM-bound kernel from a .ncu-rep (read on the CPU box)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in want:
        if k in d: print(f"{k:75s} {d[k]} {units[hdr.index(k)]}")
    st = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): float(v) for h, v in d.items()
          if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h and v}
    print("stalls per issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    print("-" * 60)
