#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            print(f"{k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
    st = []
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("stalled warps per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    print("-" * 100)
