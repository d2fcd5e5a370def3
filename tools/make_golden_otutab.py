#!/usr/bin/env python3
"""Golden files for -otutab from the UNMODIFIED reference binary (oracle/_ref/usearch12, -threads 1):
tests/golden/otutab_{reads,otus}.fa.gz (the golden reads / database relabelled with sample, size and
otu annotations in every form label.cpp understands) -> otutab.tab.gz, otutab.map.gz."""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "usearch12")
G = os.path.join(ROOT, "tests", "golden")
ql, qs = util.read_fasta(os.path.join(G, "q.fa.gz"))
dl, ds = util.read_fasta(os.path.join(G, "db.fa.gz"))
reads = []
for i, s in enumerate(qs):
    k = i % 7
    form = i % 4
    if form == 0:
        lab = "S%d.%d;size=%d;" % (k, i, 1 + i % 3)
    elif form == 1:
        lab = "r%d;sample=X%d;" % (i, k)
    elif form == 2:
        lab = "r%d;barcodelabel=B%d;size=2" % (i, k % 3)
    else:
        lab = "P%d-%d" % (k % 2, i)
    reads.append((lab, s))
otus = []
for i, s in enumerate(ds):
    form = i % 3
    if form == 0:
        lab = "Otu%d" % i
    elif form == 1:
        lab = "db%d;otu=Zotu%d;" % (i, i % 40)
    else:
        lab = "acc%d|x desc;size=5;" % i
    otus.append((lab, s))


def write(path, recs):
    with gzip.open(path, "wt") as f:
        for lab, s in recs:
            f.write(">%s\n%s\n" % (lab, s))


write(os.path.join(G, "otutab_reads.fa.gz"), reads)
write(os.path.join(G, "otutab_otus.fa.gz"), otus)
with tempfile.TemporaryDirectory() as tmp:
    for n in ("otutab_reads", "otutab_otus"):
        with gzip.open(os.path.join(G, n + ".fa.gz"), "rb") as f, open(os.path.join(tmp, n + ".fa"), "wb") as g:
            g.write(f.read())
    subprocess.run([REF, "-otutab", os.path.join(tmp, "otutab_reads.fa"), "-otus", os.path.join(tmp, "otutab_otus.fa"),
                    "-otutabout", os.path.join(tmp, "tab.txt"), "-mapout", os.path.join(tmp, "map.txt"), "-threads", "1",
                    "-quiet"], check=True)
    for src, dst in (("tab.txt", "otutab.tab.gz"), ("map.txt", "otutab.map.gz")):
        with open(os.path.join(tmp, src), "rb") as f, gzip.open(os.path.join(G, dst), "wb") as g:
            g.write(f.read())
    print(open(os.path.join(tmp, "tab.txt")).read()[:600])
    print(sum(1 for _ in open(os.path.join(tmp, "map.txt"))), "mapped reads")
