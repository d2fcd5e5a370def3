#!/usr/bin/env python3
"""Per-source-line summary of an ncu report: tools/ncu_lines.py REPORT.ncu-rep [top N]
Prints, for the hottest CUDA source lines, warp instructions executed, average active threads,
stall samples and shared-memory wavefront excess (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr[2:], r[2:]))
    try:
        lines.append((cur_file, int(r[0]), r[1].strip(), float(d["Instructions Executed"]),
                      float(d["Thread Instructions Executed"]), float(d["# Samples"]),
                      float(d.get("L1 Wavefronts Shared", 0) or 0), float(d.get("L1 Wavefronts Shared Ideal", 0) or 0),
                      float(d.get("stall_long_sb", 0) or 0), float(d.get("stall_short_sb", 0) or 0),
                      float(d.get("stall_wait", 0) or 0)))
    except (KeyError, ValueError):
        continue
tot_i = sum(l[3] for l in lines) or 1
tot_s = sum(l[5] for l in lines) or 1
print("total warp insts %.3g, samples %d" % (tot_i, tot_s))
print("%-18s %5s %6s %6s %5s %7s %6s %6s  %s" % ("file", "line", "inst%", "samp%", "thr", "smemX", "longsb", "shortsb", "source"))
for l in sorted(lines, key=lambda x: -x[5])[:top]:
    thr = l[4] / l[3] if l[3] else 0
    print("%-18s %5d %6.2f %6.2f %5.1f %7.2f %6.0f %6.0f  %s" % (l[0][:18], l[1], 100 * l[3] / tot_i, 100 * l[5] / tot_s, thr,
                                                        (l[6] / l[7]) if l[7] else 0, l[8], l[9], l[2][:90]))

# optional: per-function share (ranges given as name:lo-hi after the top-N argument)
if len(sys.argv) > 3:
    for spec in sys.argv[3:]:
        name, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        sel = [l for l in lines if l[0].startswith("usb_align") and lo <= l[1] <= hi]
        print("%-14s inst %.1f%%  samples %.1f%%  avg thr %.1f" % (name, 100 * sum(l[3] for l in sel) / tot_i,
              100 * sum(l[5] for l in sel) / tot_s, sum(l[4] for l in sel) / max(1, sum(l[3] for l in sel))))
