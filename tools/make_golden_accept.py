#!/usr/bin/env python3
"""Golden fixtures for the Accepter / Terminator / HitMgr options of -usearch_global
(accepter.cpp:41-94,145-197; terminator.cpp:66-86; hitmgr.cpp:367-398; outputsink.cpp:392):
outputs of the UNMODIFIED reference binary (oracle/_ref/usearch12), one variant per rule group.

Inputs (tests/golden/acc_db.fa.gz, acc_q.fa.gz): the targets and a slice of the reads of the main
golden set with ;size= annotations, plus queries that are copies of targets (same label and/or
same letters) for -self / -notself / -selfid.

Only runs where the reference binary exists.   Usage: python tools/make_golden_accept.py
"""
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
from tests import util  # noqa: E402

REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"

VARIANTS = {
    "acc_self": ["-id", "0.9", "-strand", "plus", "-self", "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_notself": ["-id", "0.9", "-strand", "plus", "-notself", "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_selfid": ["-id", "0.9", "-strand", "both", "-selfid", "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_cov": ["-id", "0.8", "-strand", "plus", "-query_cov", "0.95", "-target_cov", "0.17", "-max_target_cov", "0.25",
                "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_maxqcov": ["-id", "0.8", "-strand", "plus", "-max_query_cov", "0.99", "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_cols": ["-id", "0.8", "-strand", "both", "-maxid", "0.98", "-mincols", "200", "-maxgaps", "3", "-maxdiffs", "20",
                 "-mindiffs", "1", "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_size": ["-id", "0.8", "-strand", "plus", "-min_sizeratio", "0.5", "-abskew", "2.0", "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_qt": ["-id", "0.8", "-strand", "plus", "-minqt", "0.15", "-maxqt", "0.9", "-minsl", "0.17", "-maxsl", "0.95",
               "-maxaccepts", "4", "-maxrejects", "32"],
    "acc_termid": ["-id", "0.8", "-strand", "plus", "-maxaccepts", "8", "-maxrejects", "64", "-termid", "0.9"],
    "acc_termidd": ["-id", "0.8", "-strand", "plus", "-maxaccepts", "8", "-maxrejects", "64", "-termidd", "0.05"],
    "acc_maxhits": ["-id", "0.8", "-strand", "both", "-maxaccepts", "8", "-maxrejects", "64", "-maxhits", "2"],
    "acc_tophit": ["-id", "0.8", "-strand", "both", "-maxaccepts", "8", "-maxrejects", "64", "-top_hit_only"],
    "acc_tophits": ["-id", "0.8", "-strand", "both", "-maxaccepts", "8", "-maxrejects", "64", "-top_hits_only"],
    "acc_nohits": ["-id", "0.97", "-strand", "plus", "-output_no_hits"],
}


# The same rules on the other candidate loops: -usearch_local (nucleotide and protein) and amino acid
# -usearch_global.  (With -usearch_local the reference binary dies with SIGSEGV -- or survives with
# corrupted results -- as soon as a RejectPair rule rejects a pair: -self, -notself, -selfid,
# -min_sizeratio, -minqt/-maxqt, -minsl/-maxsl.  No golden files for those; the library refuses them in
# local mode.)  name -> (command, query file, database file, options); inputs are existing fixtures.
VARIANTS2 = {
    "accl_nt_cov": ("-usearch_local", "loc_nt_q.fa.gz", "loc_nt_db.fa.gz",
                    ["-id", "0.8", "-evalue", "1e-3", "-strand", "both", "-maxaccepts", "3", "-maxrejects", "8", "-query_cov", "0.9",
                     "-maxgaps", "2"]),
    "accl_nt_skew": ("-usearch_local", "acc_q.fa.gz", "acc_db.fa.gz",
                     ["-id", "0.8", "-evalue", "1e-3", "-strand", "plus", "-maxaccepts", "2", "-maxrejects", "8", "-abskew", "2.0",
                      "-max_target_cov", "0.9"]),
    "accl_aa_tcov": ("-usearch_local", "loc_aa_q.fa.gz", "loc_aa_db.fa.gz",
                     ["-id", "0.3", "-evalue", "10", "-maxaccepts", "4", "-maxrejects", "64", "-target_cov", "0.5", "-mincols", "100",
                      "-maxdiffs", "120"]),
    "accl_aa_maxid": ("-usearch_local", "loc_aa_q.fa.gz", "loc_aa_db.fa.gz",
                      ["-id", "0.5", "-evalue", "1e-5", "-maxaccepts", "2", "-maxrejects", "8", "-maxid", "0.97",
                       "-mindiffs", "2", "-max_query_cov", "0.99"]),
    "accg_aa_diffs": ("-usearch_global", "loc_aa_q.fa.gz", "loc_aa_db.fa.gz",
                      ["-id", "0.5", "-maxaccepts", "4", "-maxrejects", "16", "-maxid", "0.95", "-mindiffs", "3", "-query_cov", "0.8"]),
    "accg_aa_qt": ("-usearch_global", "loc_aa_q.fa.gz", "loc_aa_db.fa.gz",
                   ["-id", "0.3", "-maxaccepts", "8", "-maxrejects", "32", "-minqt", "0.8", "-maxqt", "1.2", "-selfid"]),
}


def build_inputs():
    g = util.Golden()
    rng = random.Random(20261020)
    db = [("db%d;size=%d" % (i, rng.choice([1, 2, 3, 5, 8, 20, 100])), s) for i, s in enumerate(g.db)]
    qs = []
    for i in list(range(0, 500)) + list(range(2400, len(g.q))):
        qs.append(("%s;size=%d" % (g.q_labels[i].split(";")[0] + "_%d" % i, rng.choice([1, 2, 4, 10, 50])), g.q[i]))
    for k in range(40):        # copies of targets: same label and letters / same label, mutated / other label, same letters
        t = rng.randrange(len(db))
        lab, s = db[t]
        if k % 3 == 0:
            qs.append((lab, s))
        elif k % 3 == 1:
            qs.append((lab, util.mutate(s, 0.02, rng)))
        else:
            qs.append(("copy%d;size=%d" % (k, rng.choice([1, 3, 9])), s))
    return db, qs


def write_fa(path, recs):
    with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as f:
        for lab, s in recs:
            f.write((">%s\n" % lab).encode())
            for i in range(0, len(s), 80):
                f.write((s[i:i + 80] + "\n").encode())


def main():
    db, qs = build_inputs()
    write_fa(os.path.join(OUT, "acc_db.fa.gz"), db)
    write_fa(os.path.join(OUT, "acc_q.fa.gz"), qs)
    tmp = tempfile.mkdtemp()
    try:
        for name in ("db", "q"):
            with gzip.open(os.path.join(OUT, "acc_%s.fa.gz" % name), "rb") as fi, open(os.path.join(tmp, name + ".fa"), "wb") as fo:
                fo.write(fi.read())
        for name, extra in VARIANTS.items():
            cmd = [REF, "-usearch_global", os.path.join(tmp, "q.fa"), "-db", os.path.join(tmp, "db.fa"), "-threads", "1", "-quiet",
                   "-uc", os.path.join(tmp, "uc"), "-blast6out", os.path.join(tmp, "b6"), "-userout", os.path.join(tmp, "user"),
                   "-userfields", USERFIELDS] + extra
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            sizes = []
            for x in ("user", "uc", "b6"):
                data = open(os.path.join(tmp, x), "rb").read()
                sizes.append(data.count(b"\n"))
                with gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (name, x)), "wb", compresslevel=9, mtime=0) as fo:
                    fo.write(data)
            print("golden", name, sizes)
        for name, (cmdname, qf, df, extra) in VARIANTS2.items():
            for src, dst in ((qf, "q2.fa"), (df, "d2.fa")):
                with gzip.open(os.path.join(OUT, src), "rb") as fi, open(os.path.join(tmp, dst), "wb") as fo:
                    fo.write(fi.read())
            cmd = [REF, cmdname, os.path.join(tmp, "q2.fa"), "-db", os.path.join(tmp, "d2.fa"), "-threads", "1", "-quiet",
                   "-uc", os.path.join(tmp, "uc"), "-blast6out", os.path.join(tmp, "b6"), "-userout", os.path.join(tmp, "user"),
                   "-userfields", USERFIELDS] + extra
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            sizes = []
            for x in ("user", "uc", "b6"):
                data = open(os.path.join(tmp, x), "rb").read()
                sizes.append(data.count(b"\n"))
                with gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (name, x)), "wb", compresslevel=9, mtime=0) as fo:
                    fo.write(data)
            print("golden", name, sizes)
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
