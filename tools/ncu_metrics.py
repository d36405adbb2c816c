#!/usr/bin/env python
"""Prints selected metrics from an .ncu-rep (raw page): python tools/ncu_metrics.py rep [substr ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_issued.avg.pct", "sm__warps_active.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct",
    "issue_stalled", "launch__registers_per_thread ", "smsp__inst_executed.sum ", "sm__inst_executed.sum ", "launch__grid_size", "lts__t_sector_hit_rate", "smsp__inst_executed_op_ldgsts", "sm__inst_executed_pipe_uniform", "inst_executed_pipe_alu", "inst_executed_pipe_fma"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        hh = h + " "
        if any(k in hh for k in keys):
            print(f"  {h} [{u}] = {v}")
