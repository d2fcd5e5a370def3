#!/usr/bin/env python3
"""Golden fixtures for exhaustive searches (tests/golden/exh_*): -maxaccepts 0 and/or -maxrejects 0
(terminator.cpp:23-31) against a database of 3 000 targets, where the candidate list of a query is longer than
the 1 024 candidates the U-sort kernel materialises.  Written by the UNMODIFIED reference binary
(oracle/_ref/usearch12); the inputs are regenerated from a seed (tools/gen_synth.py), only the outputs are
stored.   Usage: python tools/make_golden_exhaustive.py"""
import gzip
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(HERE, "tools"))
from gen_synth import generate  # noqa: E402

REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"
# name -> (options, usb_params / oracle fields)
VARIANTS = {
    "exh_00": (["-id", "0.9", "-strand", "both", "-maxaccepts", "0", "-maxrejects", "0"],
               dict(id=0.9, strand_both=1, maxaccepts=0, maxrejects=0)),
    "exh_0r": (["-id", "0.93", "-strand", "plus", "-maxaccepts", "0", "-maxrejects", "24"],
               dict(id=0.93, strand_both=0, maxaccepts=0, maxrejects=24)),
    "exh_a0": (["-id", "0.95", "-strand", "plus", "-maxaccepts", "5", "-maxrejects", "0"],
               dict(id=0.95, strand_both=0, maxaccepts=5, maxrejects=0)),
}


# -usearch_local with the same inputs: LocalAligner2 against every candidate of the list
LOCAL_USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand"
LOCAL = ("exh_loc", ["-id", "0.93", "-evalue", "1e-3", "-strand", "plus", "-maxaccepts", "0", "-maxrejects", "0"],
         dict(id=0.93, maxaccepts=0, maxrejects=0), 1e-3)


# amino acid -usearch_global (k_align): 1 500 proteins in 30 families, 150 queries
AMINO = ("exh_aa", ["-id", "0.6", "-maxaccepts", "0", "-maxrejects", "0"], dict(id=0.6, maxaccepts=0, maxrejects=0))


# RejectPair rules skip pairs without a Terminator call (searcher.cpp:63-67): with -minqt 0.5 every pair of these
# inputs is skipped (200 / 600 letters), so a query walks its whole list -- more than the 1 024 candidates of
# k_rank for two of them; with -minqt 0.335 about half of the targets are skipped
SKIPS = {
    "exh_qt5": ["-id", "0.9", "-strand", "plus", "-maxaccepts", "2", "-maxrejects", "8", "-minqt", "0.5"],
    "exh_qt3": ["-id", "0.9", "-strand", "plus", "-maxaccepts", "2", "-maxrejects", "8", "-minqt", "0.335"],
}


def inputs_aa():
    import gen_synth_aa
    db, qs = gen_synth_aa.generate(ndb=1500, length=300, nq=150, seed=31, nroot=30)
    return db, ["p%d" % i for i in range(len(db))], [r[1] for r in qs], [r[0] for r in qs]


def inputs():
    """3 000 targets of 600 letters in 30 families, 240 reads of 200 letters (12 of them random)."""
    db, reads = generate(ndb=3000, dblen=600, nq=240, qlen=200, seed=29, nroot=30)
    return db, ["db%d" % i for i in range(len(db))], [r[1] for r in reads], [r[0][1:] for r in reads]


def main():
    db, dlab, qs, qlab = inputs()
    with tempfile.TemporaryDirectory() as tmp:
        q, d = os.path.join(tmp, "q.fa"), os.path.join(tmp, "db.fa")
        open(q, "w").write("".join(">%s\n%s\n" % x for x in zip(qlab, qs)))
        open(d, "w").write("".join(">%s\n%s\n" % x for x in zip(dlab, db)))
        runs = [(name, "-usearch_global", opts, USERFIELDS) for name, (opts, _) in VARIANTS.items()]
        runs.append((LOCAL[0], "-usearch_local", LOCAL[1], LOCAL_USERFIELDS))
        runs += [(name, "-usearch_global", opts, USERFIELDS) for name, opts in SKIPS.items()]
        for name, cmd, opts, fields in runs:
            outs = {k: os.path.join(tmp, "o." + k) for k in ("user", "uc")}
            subprocess.run([REF, cmd, q, "-db", d, "-threads", "1", "-quiet"] + opts + [
                "-userout", outs["user"], "-userfields", fields, "-uc", outs["uc"]], check=True,
                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for k, path in outs.items():
                data = open(path, "rb").read()
                with gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (name, k)), "wb", compresslevel=9, mtime=0) as f:
                    f.write(data)
                print("golden", name, k, data.count(b"\n"), "lines", os.path.getsize(os.path.join(OUT, "%s.%s.gz" % (name, k))))
    db, dlab, qs, qlab = inputs_aa()
    with tempfile.TemporaryDirectory() as tmp:
        q, d = os.path.join(tmp, "q.fa"), os.path.join(tmp, "db.fa")
        open(q, "w").write("".join(">%s\n%s\n" % x for x in zip(qlab, qs)))
        open(d, "w").write("".join(">%s\n%s\n" % x for x in zip(dlab, db)))
        outs = {k: os.path.join(tmp, "o." + k) for k in ("user", "uc")}
        subprocess.run([REF, "-usearch_global", q, "-db", d, "-threads", "1", "-quiet"] + AMINO[1] + [
            "-userout", outs["user"], "-userfields", USERFIELDS, "-uc", outs["uc"]], check=True,
            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for k, path in outs.items():
            data = open(path, "rb").read()
            with gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (AMINO[0], k)), "wb", compresslevel=9, mtime=0) as f:
                f.write(data)
            print("golden", AMINO[0], k, data.count(b"\n"), "lines", os.path.getsize(os.path.join(OUT, "%s.%s.gz" % (AMINO[0], k))))


if __name__ == "__main__":
    sys.exit(main())
