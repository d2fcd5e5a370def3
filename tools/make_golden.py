#!/usr/bin/env python3
"""Builds the golden fixtures under tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/usearch12, see oracle/Makefile.ref) on small deterministic inputs.

Only runs where the reference binary exists (the build container); the outputs are committed so
that tests on the GPU box never need /root/reference.   Usage: python tools/make_golden.py
"""
import gzip
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(HERE, "tools"))
from gen_synth import generate, mutate  # noqa: E402

REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(COMP.get(c.upper(), c) for c in reversed(s))


def build_inputs():
    rng = random.Random(20261017)
    db, reads = generate(ndb=400, dblen=1500, nq=2400, qlen=250, seed=11, nroot=8)
    # --- DB edge cases: masking triggers, wildcards, short/odd lengths
    db[3] = db[3][:200] + "A" * 23 + db[3][223:]                 # homopolymer run -> fastnucleo mask
    db[5] = db[5][:400] + "AC" * 14 + db[5][428:]                # dinucleotide repeat
    db[7] = db[7][:100] + "N" * 6 + db[7][106:900] + "R" + db[7][901:]  # wildcards in target
    db[9] = db[9][:700].lower() + db[9][700:]                    # lower-case input (upper-cased by mask)
    db.append("ACGTACGTA")                                       # shorter than 2*hspw
    db.append(db[11][300:560])                                   # short target (query longer than target)
    db.append(mutate(db[13][0:900], 0.02, rng))                  # medium target
    reads = list(reads)
    n0 = len(reads)

    def add(tag, s):
        reads.append((">q%d;%s" % (len(reads), tag), s))

    for k in range(120):                                         # minus-strand reads
        t = rng.randrange(400)
        p = rng.randrange(0, 1100)
        add("rc;t=db%d" % t, revcomp(mutate(db[t][p:p + 250].upper(), rng.uniform(0, 0.03), rng)))
    for k in range(60):                                          # reads with Ns / IUPAC / lower case
        t = rng.randrange(400)
        p = rng.randrange(0, 1100)
        s = list(mutate(db[t][p:p + 250].upper(), rng.uniform(0, 0.02), rng))
        for _ in range(rng.randrange(1, 6)):
            s[rng.randrange(len(s))] = rng.choice("NNNRYKMacgtn")
        if k % 5 == 0:
            a = rng.randrange(0, 200)
            s[a:a + 30] = [c.lower() for c in s[a:a + 30]]
        add("wild;t=db%d" % t, "".join(s))
    for L in (5, 7, 8, 9, 12, 16, 20, 31, 40, 63, 64, 65, 66, 100, 128):   # short reads
        t = rng.randrange(400)
        add("short%d;t=db%d" % (L, t), db[t][500:500 + L].upper())
    for k in range(40):                                          # long reads (LA ~ LB, LA > LB)
        t = rng.randrange(400)
        s = mutate(db[t].upper(), rng.uniform(0, 0.03), rng)
        if k % 4 == 0:
            s = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 120))) + s
        if k % 4 == 1:
            s = s + "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 120)))
        if k % 4 == 2:
            s = s[rng.randrange(1, 200):]
        add("long;t=db%d" % t, s)
    for k in range(40):                                          # reads with internal indels (gapped holes)
        t = rng.randrange(400)
        p = rng.randrange(0, 1000)
        s = db[t][p:p + 320].upper()
        a = rng.randrange(60, 200)
        g = rng.randrange(1, 12)
        s = s[:a] + (s[a + g:] if k % 2 else "".join(rng.choice("ACGT") for _ in range(g)) + s[a:])
        add("indel;t=db%d" % t, mutate(s, 0.01, rng))
    for k in range(20):                                          # exact copies and low-complexity queries
        t = rng.randrange(400)
        add("exact;t=db%d" % t, db[t][100:350].upper())
    add("polyA", "A" * 250)
    add("dinuc", "AC" * 125)
    add("allN", "N" * 100)
    return db, reads, n0


def write_fa(path, recs):
    with gzip.open(path, "wt", compresslevel=9) as f:
        for lab, s in recs:
            f.write("%s\n%s\n" % (lab, s))


def run(name, q, d, extra):
    tmp = os.path.join(OUT, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    cmd = [REF, "-usearch_global", q, "-db", d, "-threads", "1", "-quiet",
           "-uc", os.path.join(tmp, "uc"), "-blast6out", os.path.join(tmp, "b6"),
           "-userout", os.path.join(tmp, "user"), "-userfields", USERFIELDS] + extra
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for x in ("user", "uc", "b6"):
        # -threads 1 => file order is query input order, hits in HitMgr order: keep it
        lines = open(os.path.join(tmp, x)).read().splitlines()
        with gzip.open(os.path.join(OUT, "%s.%s.gz" % (name, x)), "wt", compresslevel=9) as f:
            f.write("\n".join(lines) + "\n")
        os.remove(os.path.join(tmp, x))
    os.rmdir(tmp)


def main():
    os.makedirs(OUT, exist_ok=True)
    db, reads, _ = build_inputs()
    dpath, qpath = os.path.join(OUT, "db.fa.gz"), os.path.join(OUT, "q.fa.gz")
    write_fa(dpath, [(">db%d" % i, s) for i, s in enumerate(db)])
    write_fa(qpath, reads)
    # the reference reads .gz directly (gzipfileio.cpp) but keep it simple: plain temp copies
    dfa, qfa = os.path.join(OUT, "_db.fa"), os.path.join(OUT, "_q.fa")
    for src, dst in ((dpath, dfa), (qpath, qfa)):
        with gzip.open(src, "rt") as fi, open(dst, "w") as fo:
            fo.write(fi.read())
    variants = {
        "plus97": ["-id", "0.97", "-strand", "plus"],
        "both97": ["-id", "0.97", "-strand", "both"],
        "plus90_ma4": ["-id", "0.9", "-strand", "plus", "-maxaccepts", "4", "-maxrejects", "64"],
        "both80_ma0": ["-id", "0.8", "-strand", "both", "-maxaccepts", "3", "-maxrejects", "16"],
    }
    for name, extra in variants.items():
        run(name, qfa, dfa, extra)
        print("golden", name)
    # extended userfields (first/last-M coordinates, gap counts, full path): user file only
    xf = ("query+target+id+fractid+dist+pairs+gaps+allgaps+qlot+qhit+qunt+tlot+thit+tunt+ql+tl+alnlen+opens+exts+"
          "aln+tstrand+mism+ids+diffs+clusternr")
    tmp = os.path.join(OUT, "_x.user")
    subprocess.run([REF, "-usearch_global", qfa, "-db", dfa, "-id", "0.9", "-strand", "both", "-maxaccepts", "2",
                    "-maxrejects", "16", "-threads", "1", "-quiet", "-userout", tmp, "-userfields", xf], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with gzip.open(os.path.join(OUT, "both90x.user.gz"), "wt", compresslevel=9) as f:
        f.write(open(tmp).read())
    os.remove(tmp)
    os.remove(dfa)
    os.remove(qfa)


if __name__ == "__main__":
    main()
