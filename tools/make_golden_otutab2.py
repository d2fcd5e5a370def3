#!/usr/bin/env python3
"""More golden files for -otutab from the UNMODIFIED reference binary (-threads 1) on the inputs of
tools/make_golden_otutab.py: otutab.biom.gz (-biomout, json.cpp:32-110; the "date" line is blanked) and
otutab.hits.gz (its -userout with the fields of a usb_hit: the hit table tools/format_replay.cpp feeds to the
OtuTabSink on a machine without a GPU).   Usage: python tools/make_golden_otutab2.py"""
import gzip
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden_formats as M  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def blank_date(data):
    return re.sub(rb'"date": "[^"]*"', b'"date": ""', data)


def main():
    with tempfile.TemporaryDirectory() as tmp:
        for n in ("otutab_reads", "otutab_otus"):
            with gzip.open(os.path.join(G, n + ".fa.gz"), "rb") as f, open(os.path.join(tmp, n + ".fa"), "wb") as g:
                g.write(f.read())
        subprocess.run([M.REF, "-otutab", "otutab_reads.fa", "-otus", "otutab_otus.fa", "-otutabout", "tab.txt", "-mapout",
                        "map.txt", "-biomout", "o.biom", "-dbmatched", "dbm.fa", "-dbnotmatched", "dbnm.fa", "-notmatched", "nm.fa", "-sizeout",
                        "-userout", "hits.txt", "-userfields", M.HITFIELDS, "-threads", "1",
                        "-quiet"], check=True, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for src, dst in (("tab.txt", "otutab.tab.gz"), ("map.txt", "otutab.map.gz")):  # unchanged by the extra outputs
            with gzip.open(os.path.join(G, dst), "rb") as g:
                assert g.read() == open(os.path.join(tmp, src), "rb").read(), dst
        for src, dst, fix in (("o.biom", "otutab.biom.gz", blank_date), ("hits.txt", "otutab.hits.gz", lambda d: d)):
            data = fix(open(os.path.join(tmp, src), "rb").read())
            with gzip.GzipFile(os.path.join(G, dst), "wb", compresslevel=9, mtime=0) as g:
                g.write(data)
            print(dst, data.count(b"\n"), "lines")
        # DBHitSink counts one hit per query in -otutab (dbhitsink.cpp:138-139); digests of the three FASTA files
        import hashlib
        import json
        sums = {k: {"sha256": hashlib.sha256(open(os.path.join(tmp, k), "rb").read()).hexdigest(),
                    "seqs": open(os.path.join(tmp, k), "rb").read().count(b">")} for k in ("dbm.fa", "dbnm.fa", "nm.fa")}
        with open(os.path.join(G, "otutab_sha256.json"), "w") as f:
            json.dump(sums, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    sys.exit(main())
