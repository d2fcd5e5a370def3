#!/usr/bin/env python3
"""tests/golden/acc_sel.hits.gz: the hit table (reference -userout with the fields of a usb_hit) of the search behind
the goldens acc_maxhits / acc_tophit / acc_tophits (tools/make_golden_accept.py) WITHOUT the selection option, so that
tools/format_replay.cpp can apply -maxhits / -top_hit_only / -top_hits_only (hitmgr.cpp:367-420,466-475) on the host
and the result be compared with those goldens on a machine without a GPU.   Usage: python tools/make_golden_selection.py"""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden_formats as M  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
OPTS = ["-id", "0.8", "-strand", "both", "-maxaccepts", "8", "-maxrejects", "64"]
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"
SELECTIONS = {"acc_maxhits": ["-maxhits", "2"], "acc_tophit": ["-top_hit_only"], "acc_tophits": ["-top_hits_only"]}


def main():
    with tempfile.TemporaryDirectory() as tmp:
        for n in ("acc_q", "acc_db"):
            with gzip.open(os.path.join(G, n + ".fa.gz"), "rb") as f, open(os.path.join(tmp, n + ".fa"), "wb") as g:
                g.write(f.read())
        subprocess.run([M.REF, "-usearch_global", "acc_q.fa", "-db", "acc_db.fa", "-threads", "1", "-quiet"] + OPTS + [
            "-userout", "hits.txt", "-userfields", M.HITFIELDS], check=True, cwd=tmp, stdout=subprocess.DEVNULL,
            stderr=subprocess.DEVNULL)
        data = open(os.path.join(tmp, "hits.txt"), "rb").read()
        with gzip.GzipFile(os.path.join(G, "acc_sel.hits.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(data)
        print("acc_sel.hits.gz", data.count(b"\n"), "lines")


if __name__ == "__main__":
    sys.exit(main())
