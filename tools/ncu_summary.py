#!/usr/bin/env python3
"""Kernel-level summary of an ncu report: tools/ncu_summary.py REPORT.ncu-rep [launch index]
Duration, instructions, issue utilisation, occupancy, DRAM bytes and the stall mix."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("== kernel:", r[h.index("Kernel Name")][:60] if "Kernel Name" in h else "")
    for i, n in enumerate(h):
        if n in want:
            print("  %-70s %s %s" % (n, r[i], rows[1][i]))
    st = [(float(r[i]), n) for i, n in enumerate(h) if "average_warps_issue_stalled" in n and n.endswith("per_issue_active.ratio") and r[i]]
    for v, n in sorted(st, reverse=True)[:8]:
        print("  stall %-40s %.3f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
