#!/usr/bin/env python3
"""Golden fixtures for -usearch_local on sequences longer than g_MaxL = 4096 letters (xdpmem.h:6), where
the reference splits X-drop extensions (xdropfwdsplit.cpp, xdropbwdsplit.cpp): tests/golden/loclong_*,
outputs of the UNMODIFIED reference binary (oracle/_ref/usearch12, -threads 1).
Usage: python tools/make_golden_local_long.py"""
import gzip
import os
import random
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand"
AA = "ACDEFGHIKLMNPQRSTVWY"


def mut(s, r, alpha, rng):
    out = []
    for c in s:
        x = rng.random()
        if x < r * 0.6:
            out.append(rng.choice(alpha))
        elif x < r * 0.8:
            continue
        elif x < r:
            out.append(c)
            out.append(rng.choice(alpha))
        else:
            out.append(c)
    return "".join(out)


def build_nt():
    rng = random.Random(20261101)
    db = ["".join(rng.choice("ACGT") for _ in range(rng.choice([3000, 5000, 9000, 13000, 20000]))) for _ in range(24)]
    qs = []
    for i in range(48):
        t = rng.randrange(len(db))
        s = db[t]
        a = rng.randrange(0, max(1, len(s) - 3000))
        b = min(len(s), a + rng.choice([1500, 4090, 4097, 8192, 8200, 12000, 19000]))
        q = mut(s[a:b], rng.choice([0.02, 0.05, 0.1, 0.18]), "ACGT", rng)
        if rng.random() < 0.4:   # an insertion the X-drop may or may not bridge
            k = rng.randrange(len(q))
            q = q[:k] + "".join(rng.choice("ACGT") for _ in range(rng.choice([30, 300]))) + q[k:]
        if i % 6 == 5:           # minus strand
            q = q[::-1].translate(str.maketrans("ACGT", "TGCA"))
        qs.append((">q%d;t=%d" % (i, t), q))
    return [(">t%d" % i, s) for i, s in enumerate(db)], qs


def build_aa():
    rng = random.Random(20261102)
    db = ["".join(rng.choice(AA) for _ in range(rng.choice([800, 4500, 6000, 9000]))) for _ in range(24)]
    qs = []
    for i in range(40):
        t = rng.randrange(len(db))
        s = db[t]
        a = rng.randrange(0, max(1, len(s) - 500))
        b = min(len(s), a + rng.choice([400, 4200, 5000, 8800]))
        qs.append((">q%d;t=%d" % (i, t), mut(s[a:b], rng.choice([0.05, 0.15, 0.3]), AA, rng)))
    return [(">p%d" % i, s) for i, s in enumerate(db)], qs


def write_fa(path, recs):
    with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as f:
        for lab, s in recs:
            f.write(("%s\n%s\n" % (lab, s)).encode())


def run(name, q, d, extra):
    tmp = os.path.join(OUT, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    for src, dst in ((q, "q.fa"), (d, "d.fa")):
        with gzip.open(src, "rb") as fi, open(os.path.join(tmp, dst), "wb") as fo:
            fo.write(fi.read())
    cmd = [REF, "-usearch_local", os.path.join(tmp, "q.fa"), "-db", os.path.join(tmp, "d.fa"), "-threads", "1", "-quiet",
           "-uc", os.path.join(tmp, "uc"), "-blast6out", os.path.join(tmp, "b6"),
           "-userout", os.path.join(tmp, "user"), "-userfields", USERFIELDS] + extra
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for x in ("user", "uc", "b6"):
        with open(os.path.join(tmp, x), "rb") as fi, gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (name, x)), "wb", mtime=0) as fo:
            fo.write(fi.read())
        print(name, x, sum(1 for _ in open(os.path.join(tmp, x))), "lines")
    import shutil
    shutil.rmtree(tmp)


if __name__ == "__main__":
    db, qs = build_nt()
    write_fa(os.path.join(OUT, "loclong_nt_db.fa.gz"), db)
    write_fa(os.path.join(OUT, "loclong_nt_q.fa.gz"), qs)
    run("loclong_nt_both", os.path.join(OUT, "loclong_nt_q.fa.gz"), os.path.join(OUT, "loclong_nt_db.fa.gz"),
        ["-id", "0.7", "-evalue", "1e-5", "-strand", "both", "-maxaccepts", "2", "-maxrejects", "16"])
    db, qs = build_aa()
    write_fa(os.path.join(OUT, "loclong_aa_db.fa.gz"), db)
    write_fa(os.path.join(OUT, "loclong_aa_q.fa.gz"), qs)
    run("loclong_aa_e5", os.path.join(OUT, "loclong_aa_q.fa.gz"), os.path.join(OUT, "loclong_aa_db.fa.gz"),
        ["-id", "0.5", "-evalue", "1e-5"])
