#!/usr/bin/env python3
"""Golden files for -fastx_uniques from the UNMODIFIED reference binary (-threads 1):
tests/golden/uniq_in.fa.gz (the golden cluster reads plus duplicates in other case and with size= /
other annotations, shuffled) -> uniq_sizeout.fa.gz, uniq_relabel.fa.gz, uniq_plain.fa.gz."""
import gzip
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "usearch12")
G = os.path.join(ROOT, "tests", "golden")
labels, seqs = util.read_fasta(os.path.join(G, "cluster_reads.fa.gz"))
rng = random.Random(3)
out = []
for i, (l, s) in enumerate(zip(labels, seqs)):
    out.append((l, s))
    if i % 9 == 0:
        out.append(("dup%d;size=%d;" % (i, i % 5 + 1), s.lower() if i % 2 else s))
    if i % 31 == 0:
        out.append(("x%d;size=7;foo=bar" % i, s))
rng.shuffle(out)
with gzip.open(os.path.join(G, "uniq_in.fa.gz"), "wt") as f:
    for l, s in out:
        f.write(">%s\n%s\n" % (l, s))
with tempfile.TemporaryDirectory() as tmp:
    src = os.path.join(tmp, "in.fa")
    with gzip.open(os.path.join(G, "uniq_in.fa.gz"), "rb") as f, open(src, "wb") as g:
        g.write(f.read())
    for name, extra in (("uniq_sizeout", ["-sizeout"]), ("uniq_relabel", ["-sizeout", "-relabel", "Uniq", "-minuniquesize", "2"]),
                        ("uniq_plain", [])):
        dst = os.path.join(tmp, name + ".fa")
        subprocess.run([REF, "-fastx_uniques", src, "-fastaout", dst, "-threads", "1", "-quiet"] + extra, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(dst, "rb") as f, gzip.open(os.path.join(G, name + ".fa.gz"), "wb") as g:
            g.write(f.read())
        print(name, sum(1 for x in open(dst) if x.startswith(">")), "uniques")
    # -sizein / -topn (derepresult.cpp:705-707,822-844): tests/golden/uniq2_<name>.fa.gz
    for name, extra in util.UNIQUES2_VARIANTS.items():
        dst = os.path.join(tmp, "o.fa")
        subprocess.run([REF, "-fastx_uniques", src, "-fastaout", dst, "-threads", "1", "-quiet"] + extra, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(dst, "rb") as f, gzip.GzipFile(os.path.join(G, "uniq2_%s.fa.gz" % name), "wb", mtime=0) as g:
            g.write(f.read())
        print("uniq2", name, sum(1 for x in open(dst) if x.startswith(">")), "uniques")
