#!/usr/bin/env python3
"""Golden fixtures for -usearch_local (tests/golden/loc_*): outputs of the UNMODIFIED reference
binary (oracle/_ref/usearch12) on small deterministic protein and nucleotide inputs.

Only runs where the reference binary exists (the build container); the outputs are committed so
that tests on the GPU box never need /root/reference.   Usage: python tools/make_golden_local.py
"""
import gzip
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(HERE, "tools"))
import gen_synth  # noqa: E402
import gen_synth_aa  # noqa: E402

REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand"
AA = gen_synth_aa.AA


def build_aa():
    rng = random.Random(20261018)
    db, qs = gen_synth_aa.generate(ndb=300, length=300, nq=500, seed=5, nroot=12)
    qs = [(">" + lab, s) for lab, s in qs]
    # --- DB edge cases
    db[2] = db[2][:80] + "A" * 17 + db[2][97:]                    # homopolymer -> fastamino mask
    db[4] = db[4][:50] + "KE" * 9 + db[4][68:]                    # dipeptide repeat -> mask
    db[6] = db[6][:100] + "XXBZ" + db[6][104:200] + "U" + db[6][201:]  # wildcards / non-standard letters
    db[8] = db[8][:150].lower() + db[8][150:]                     # lower case input
    db.append("ACDEF")                                            # shorter than 2 * seed word
    db.append("ACDEFG")
    db.append(db[10][40:160])                                     # fragment
    db.append(db[12][:150] + db[14][100:260])                     # two-domain chimera
    db.append(db[16][150:] + db[16][:150])                        # circular permutation -> 2 ARs per target
    db.append(db[18] + gen_synth_aa.mutate(db[18], 0.1, rng))     # tandem duplication

    def add(tag, s):
        qs.append((">a%d;%s" % (len(qs), tag), s))

    for k in range(40):                                           # two-domain queries
        t1, t2 = rng.randrange(300), rng.randrange(300)
        add("chim;t=p%d+p%d" % (t1, t2), gen_synth_aa.mutate(db[t1][:140] + db[t2][140:], rng.uniform(0.02, 0.3), rng))
    for k in range(30):                                           # permuted queries (several HSPs per target)
        t = rng.randrange(300)
        s = gen_synth_aa.mutate(db[t], rng.uniform(0.02, 0.25), rng)
        c = rng.randrange(80, 220)
        add("perm;t=p%d" % t, s[c:] + s[:c])
    for k in range(30):                                           # wildcards and lower case in queries
        t = rng.randrange(300)
        s = list(gen_synth_aa.mutate(db[t].upper(), rng.uniform(0.02, 0.2), rng))
        for _ in range(rng.randrange(1, 8)):
            s[rng.randrange(len(s))] = rng.choice("XXBZUOJxbz")
        if k % 3 == 0:
            a = rng.randrange(0, 200)
            s[a:a + 40] = [c.lower() for c in s[a:a + 40]]
        add("wild;t=p%d" % t, "".join(s))
    for L in (2, 3, 4, 5, 6, 7, 10, 16, 25, 40, 64):              # short queries
        t = rng.randrange(300)
        add("short%d;t=p%d" % (L, t), db[t][100:100 + L].upper())
    for k in range(30):                                           # fragments and extended queries
        t = rng.randrange(300)
        s = gen_synth_aa.mutate(db[t].upper(), rng.uniform(0.0, 0.3), rng)
        if k % 3 == 0:
            s = s[rng.randrange(0, 100):rng.randrange(150, 300)]
        elif k % 3 == 1:
            s = "".join(rng.choice(AA) for _ in range(rng.randrange(1, 150))) + s
        else:
            s = s + "".join(rng.choice(AA) for _ in range(rng.randrange(1, 150)))
        add("frag;t=p%d" % t, s)
    for k in range(20):                                           # long gaps
        t = rng.randrange(300)
        s = db[t].upper()
        a = rng.randrange(60, 200)
        g = rng.randrange(1, 25)
        s = s[:a] + (s[a + g:] if k % 2 else "".join(rng.choice(AA) for _ in range(g)) + s[a:])
        add("gap;t=p%d" % t, gen_synth_aa.mutate(s, 0.05, rng))
    for k in range(10):
        t = rng.randrange(300)
        add("exact;t=p%d" % t, db[t].upper())
    add("polyA", "A" * 200)
    add("dipep", "KE" * 100)
    add("allX", "X" * 100)
    add("random", "".join(rng.choice(AA) for _ in range(300)))
    return db, qs


def build_nt():
    rng = random.Random(20261019)
    db, reads = gen_synth.generate(ndb=200, dblen=1200, nq=500, qlen=250, seed=13, nroot=6)
    reads = list(reads)

    def add(tag, s):
        reads.append((">q%d;%s" % (len(reads), tag), s))

    for k in range(40):                                           # chimeric reads: two local hits
        t1, t2 = rng.randrange(200), rng.randrange(200)
        p1, p2 = rng.randrange(0, 900), rng.randrange(0, 900)
        add("chim;t=db%d+db%d" % (t1, t2), gen_synth.mutate(db[t1][p1:p1 + 130] + db[t2][p2:p2 + 130], 0.03, rng))
    for k in range(30):                                           # wildcards / lower case
        t = rng.randrange(200)
        p = rng.randrange(0, 900)
        s = list(gen_synth.mutate(db[t][p:p + 250], 0.03, rng))
        for _ in range(rng.randrange(1, 6)):
            s[rng.randrange(len(s))] = rng.choice("NNRYKMacgtn")
        add("wild;t=db%d" % t, "".join(s))
    for L in (4, 5, 9, 10, 11, 20, 40, 64, 100):
        t = rng.randrange(200)
        add("short%d;t=db%d" % (L, t), db[t][300:300 + L])
    for k in range(20):                                           # long reads with divergent flanks
        t = rng.randrange(200)
        s = "".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 200))) + gen_synth.mutate(db[t][200:900], 0.05, rng)
        add("long;t=db%d" % t, s)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for k in range(60):                                           # minus-strand reads (hits only with -strand both)
        t = rng.randrange(200)
        p = rng.randrange(0, 900)
        s = gen_synth.mutate(db[t][p:p + 250], rng.uniform(0, 0.05), rng)
        add("rc;t=db%d" % t, "".join(comp.get(c, c) for c in reversed(s)))
    add("polyA", "A" * 200)
    return db, reads


def write_fa(path, recs):
    with gzip.open(path, "wt", compresslevel=9) as f:
        for lab, s in recs:
            f.write("%s\n%s\n" % (lab, s))


def run(name, q, d, extra):
    tmp = os.path.join(OUT, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    cmd = [REF, "-usearch_local", q, "-db", d, "-threads", "1", "-quiet",
           "-uc", os.path.join(tmp, "uc"), "-blast6out", os.path.join(tmp, "b6"),
           "-userout", os.path.join(tmp, "user"), "-userfields", USERFIELDS] + extra
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for x in ("user", "uc", "b6"):
        lines = open(os.path.join(tmp, x)).read().splitlines()
        with gzip.open(os.path.join(OUT, "%s.%s.gz" % (name, x)), "wt", compresslevel=9) as f:
            f.write("\n".join(lines) + "\n")
        os.remove(os.path.join(tmp, x))
        print("golden", name, x, len(lines), "lines")
    os.rmdir(tmp)


def plain(src, dst):
    with gzip.open(src, "rt") as fi, open(dst, "w") as fo:
        fo.write(fi.read())


def main():
    os.makedirs(OUT, exist_ok=True)
    db, qs = build_aa()
    write_fa(os.path.join(OUT, "loc_aa_db.fa.gz"), [(">p%d" % i, s) for i, s in enumerate(db)])
    write_fa(os.path.join(OUT, "loc_aa_q.fa.gz"), qs)
    ndb, nqs = build_nt()
    write_fa(os.path.join(OUT, "loc_nt_db.fa.gz"), [(">db%d" % i, s) for i, s in enumerate(ndb)])
    write_fa(os.path.join(OUT, "loc_nt_q.fa.gz"), nqs)
    tmpf = {}
    for k in ("aa_db", "aa_q", "nt_db", "nt_q"):
        tmpf[k] = os.path.join(OUT, "_loc_%s.fa" % k)
        plain(os.path.join(OUT, "loc_%s.fa.gz" % k), tmpf[k])
    run("loc_aa_e5", tmpf["aa_q"], tmpf["aa_db"], ["-id", "0.5", "-evalue", "1e-5"])
    run("loc_aa_ma4", tmpf["aa_q"], tmpf["aa_db"], ["-id", "0.3", "-evalue", "10", "-maxaccepts", "4", "-maxrejects", "64"])
    run("loc_nt_plus", tmpf["nt_q"], tmpf["nt_db"], ["-id", "0.9", "-evalue", "1e-5", "-strand", "plus", "-maxaccepts", "2",
                                                   "-maxrejects", "16"])
    run("loc_nt_both", tmpf["nt_q"], tmpf["nt_db"], ["-id", "0.8", "-evalue", "1e-3", "-strand", "both", "-maxaccepts", "3",
                                                   "-maxrejects", "8"])
    for f in tmpf.values():
        os.remove(f)


if __name__ == "__main__":
    main()
