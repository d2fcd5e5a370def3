#!/usr/bin/env python3
"""profiles/roofline_traffic.json from ncu reports of one step of the default bench workload:
    tools/make_roofline_traffic.py TAG     (reads gpurun_out/prof_{rank,gate,dp}_TAG.ncu-rep)
Per kernel family: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and gpu__time_duration
summed over the launches of one step (k_gate and k_dp launch once per stage).  bench.py reports the
bytes as roofline.traffic only while its own timing of the kernel is within 5 % of gpu_time_ms."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]

    def col(name):
        i = h.index(name)
        return [float(r[i]) * UNIT[units[i]] for r in rows[2:]]
    rd, wr, ms = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    issue = col_raw(rows, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    inst = col_raw(rows, "smsp__inst_executed.sum")
    return {"launches": len(ms), "gpu_time_ms": sum(ms), "dram_bytes": sum(rd) + sum(wr), "dram_read": sum(rd), "dram_write": sum(wr),
            "warp_instructions": sum(inst), "issue_active_pct": issue}


def col_raw(rows, name):
    i = rows[0].index(name)
    return [float(r[i]) for r in rows[2:]]


def main():
    tag = sys.argv[1]
    out = {"workload": "usearch_global 1000000x250bp synthetic reads vs 100000x1500bp synthetic DB, -id 0.97 -strand plus",
           "source": "ncu --set full --clock-control none, one step (tools/gpu_profile_stage.sh %s); per-launch times under ncu are "
                     "cold-cache and serialised" % tag, "kernels": {}}
    for k in ("rank", "gate", "dp"):
        rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (k, tag))
        if os.path.exists(rep):
            out["kernels"]["k_" + k] = summarize(rep)
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
