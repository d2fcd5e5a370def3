#!/usr/bin/env python3
"""Golden files for -cluster_fast with size annotations (-sizein / -sizeout / -relabel / -minsize) and for
-cluster_smallmem,
made by the unmodified reference binary (oracle/_ref/usearch12, -threads 1):
    python tools/make_golden_cluster_sizes.py
The reads are tests/golden/cluster_reads.fa.gz with ";size=N;" appended to every label
(tests/util.py: sized_cluster_reads)."""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "usearch12")

with tempfile.TemporaryDirectory() as tmp:
    reads = os.path.join(tmp, "r.fa")
    util.sized_cluster_reads(reads)
    for name, extra in util.CLUSTER_SIZE_VARIANTS.items():
        uc, cen = os.path.join(tmp, "o.uc"), os.path.join(tmp, "o.fa")
        subprocess.run([REF, "-cluster_fast", reads, "-id", "0.97", "-uc", uc, "-centroids", cen, "-threads", "1", "-quiet"] + extra,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for src, out in ((uc, "cluster_%s.uc.gz" % name), (cen, "cluster_%s.centroids.fa.gz" % name)):
            with open(src, "rb") as f, gzip.GzipFile(os.path.join(util.GOLDEN, out), "wb", mtime=0) as g:
                g.write(f.read())
            print(out, os.path.getsize(os.path.join(util.GOLDEN, out)))
    # -cluster_smallmem (clustersmallmem.cpp): input order, no dereplication
    for name, (order, extra) in util.SMALLMEM_VARIANTS.items():
        util.smallmem_reads(reads, order)
        uc, cen = os.path.join(tmp, "o.uc"), os.path.join(tmp, "o.fa")
        subprocess.run([REF, "-cluster_smallmem", reads, "-id", "0.97", "-uc", uc, "-centroids", cen, "-threads", "1", "-quiet"] + extra,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for src, out in ((uc, "cluster_%s.uc.gz" % name), (cen, "cluster_%s.centroids.fa.gz" % name)):
            with open(src, "rb") as f, gzip.GzipFile(os.path.join(util.GOLDEN, out), "wb", mtime=0) as g:
                g.write(f.read())
            print(out, os.path.getsize(os.path.join(util.GOLDEN, out)))
