#!/usr/bin/env python3
"""cluster_fast timing (BASELINE config 3): tools/bench_cluster.py [--reads N] [--amplicon] [--ref-sample M]
Writes synthetic reads to a temp FASTA, times usearch12_b200_cli -cluster_fast end to end and the
reference binary (oracle/_ref/usearch12 -cluster_fast -threads 1; its search loop is serial) on the
first M reads, and prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1000000)
    ap.add_argument("--db", type=int, default=100000)
    ap.add_argument("--amplicon", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=50000)
    a = ap.parse_args()
    from usearch12_b200 import build
    cli = build.build_cli()
    db, db_off = synth_np.gen_db(a.db, 1500, seed=4)
    reads, r_off, _ = synth_np.gen_reads(db, db_off, a.reads, 250, seed=3000, window=(500, 750) if a.amplicon else None)
    tmp = tempfile.mkdtemp(prefix="usb_cl_")
    fa = os.path.join(tmp, "r.fa")
    synth_np.write_fasta(fa, reads, r_off, "r")
    t = time.perf_counter()
    r = subprocess.run([cli, "-cluster_fast", fa, "-id", "0.97", "-uc", os.path.join(tmp, "o.uc"), "-centroids",
                        os.path.join(tmp, "o.fa")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    ours = time.perf_counter() - t
    line = {"workload": "cluster_fast %dx250bp %s reads -id 0.97" % (a.reads, "amplicon" if a.amplicon else "window-random"),
            "ours_s": ours, "ours_reads_per_s": a.reads / ours, "ours_log": r.stdout.strip().splitlines()[-1:], "rc": r.returncode}
    ref = os.path.join(ROOT, "oracle", "_ref", "usearch12")
    if a.ref_sample and os.path.exists(ref):
        fs = os.path.join(tmp, "s.fa")
        synth_np.write_fasta(fs, reads, r_off, "r", 0, min(a.ref_sample, a.reads))
        t = time.perf_counter()
        subprocess.run([ref, "-cluster_fast", fs, "-id", "0.97", "-threads", "1", "-uc", os.path.join(tmp, "r.uc"), "-quiet"],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tr = time.perf_counter() - t
        line["ref_sample"] = min(a.ref_sample, a.reads)
        line["ref_s"] = tr
        line["ref_reads_per_s"] = min(a.ref_sample, a.reads) / tr
        # same sample through ours, for an identical-output check
        subprocess.run([cli, "-cluster_fast", fs, "-id", "0.97", "-uc", os.path.join(tmp, "s.uc"), "-quiet"], check=True)
        line["sample_uc_identical"] = open(os.path.join(tmp, "s.uc")).read() == open(os.path.join(tmp, "r.uc")).read()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
