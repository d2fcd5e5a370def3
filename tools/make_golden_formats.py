#!/usr/bin/env python3
"""Golden fixtures for the output formats of the search commands (tests/golden/fmt_*): files written
by the UNMODIFIED reference binary (oracle/_ref/usearch12) -- -alnout, -fastapairs, -qsegout, -tsegout,
-matched, -notmatched, -dbmatched, -dbnotmatched, -uc, -blast6out and -userout with every userfield
the host mirror implements -- on subsets of the existing golden inputs.

`<name>.hits` is the reference's -userout with the fields of a usb_hit (HITFIELDS): the hit table that
tools/format_replay.cpp feeds to the host sinks on a machine without a GPU.

Only runs where the reference binary exists (the build container); the outputs are committed so that
tests on the GPU box never need /root/reference.   Usage: python tools/make_golden_formats.py
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")

HITFIELDS = "query+target+qstrand+ids+mism+gaps+opens+qlot+qhit+tlot+thit+alnlen+ql+tl+raw+aln"
COMMON = ("query+target+evalue+id+fractid+dist+mid+pctpv+pctgaps+pairs+gaps+allgaps+qlo+qhi+tlo+thi+qlot+qhit+qunt+tlot+"
          "thit+tunt+pv+ql+tl+qs+ts+alnlen+opens+exts+raw+bits+aln+caln+qseq+tseq+qstrand+tstrand+qrow+trow+qrowdots+"
          "trowdots+qframe+tframe+mism+ids+qcov+tcov+diffs+diffsa+editdiffs+qlor+qhir+tlor+thir+orflo+orfhi+orfframe+"
          "kmerid+qsegf+clusternr")
LOCAL_ONLY = "+qseg+tseg+gc"

# name -> (command, query fixture, db fixture, keep(i) of the query fixture, search options, userfields)
VARIANTS = {
    "fmt_nt": ("usearch_global", "q.fa.gz", "db.fa.gz", lambda i: i < 20 or i >= 2400,
               ["-id", "0.9", "-strand", "both", "-maxaccepts", "2", "-maxrejects", "16"], COMMON),
    # queries without hits in -userout / -blast6out (userout.cpp:53-124, blast6out.cpp:82-103)
    "fmt_nh": ("usearch_global", "q.fa.gz", "db.fa.gz", lambda i: i < 20 or i >= 2400,
               ["-id", "0.97", "-strand", "plus", "-output_no_hits"],
               "query+target+id+mid+qs+ts+qrow+qcov+diffsa+qseq+tseq+ql+clusternr", ("hits", "user", "b6", "uc")),
    # -minsize: smaller queries are not searched and reach no sink (search.cpp:59-82); -uc_hitsonly (outputuc.cpp:14-15)
    "fmt_ms": ("usearch_global", "acc_q.fa.gz", "acc_db.fa.gz", lambda i: i % 4 == 1,
               ["-id", "0.9", "-strand", "plus", "-minsize", "5", "-uc_hitsonly"], "query+target+id",
               ("hits", "user", "uc", "matched", "notmatched")),
    # -rowlen / -flank (alnout.cpp:93-99, userout.cpp:216-235)
    "fmt_rl": ("usearch_global", "q.fa.gz", "db.fa.gz", lambda i: i >= 2560,
               ["-id", "0.9", "-strand", "both", "-maxaccepts", "2", "-maxrejects", "16", "-rowlen", "50", "-flank", "3"],
               "query+target+qsegf+qlo+qhi+qtrimlo+qtrimhi+qtrimseq", ("hits", "user", "aln", "trim")),
    "fmt_sz": ("usearch_global", "acc_q.fa.gz", "acc_db.fa.gz", lambda i: i % 8 == 0,
               ["-id", "0.9", "-strand", "plus", "-maxaccepts", "3", "-maxrejects", "16", "-sizein", "-sizeout"],
               "query+target+id+abskew+qcov+tcov"),
    "fmt_aag": ("usearch_global", "loc_aa_q.fa.gz", "loc_aa_db.fa.gz", lambda i: i < 10 or i >= 520,
                ["-id", "0.5", "-maxaccepts", "2", "-maxrejects", "16"], COMMON),
    "fmt_aal": ("usearch_local", "loc_aa_q.fa.gz", "loc_aa_db.fa.gz", lambda i: i < 10 or i >= 520,
                ["-id", "0.3", "-evalue", "10", "-maxaccepts", "4", "-maxrejects", "64"], COMMON + LOCAL_ONLY),
    "fmt_ntl": ("usearch_local", "loc_nt_q.fa.gz", "loc_nt_db.fa.gz", lambda i: i < 10 or i >= 520,
                ["-id", "0.8", "-evalue", "1e-3", "-strand", "both", "-maxaccepts", "3", "-maxrejects", "16"],
                COMMON + LOCAL_ONLY),
}
KINDS = ("hits", "user", "aln", "pairs", "qseg", "tseg", "matched", "notmatched", "uc", "b6")
FLAGS = {"aln": "-alnout", "pairs": "-fastapairs", "qseg": "-qsegout", "tseg": "-tsegout", "matched": "-matched",
         "notmatched": "-notmatched", "uc": "-uc", "b6": "-blast6out", "dbm": "-dbmatched", "dbnm": "-dbnotmatched",
         "dbcut": "-dbcutout", "trim": "-trimout"}


def read_fasta(path):
    recs = []
    with gzip.open(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                recs.append([line, ""])
            else:
                recs[-1][1] += line
    return recs


def kinds_of(name):
    """The stored output kinds of a variant (all of KINDS unless the variant names its own)."""
    v = VARIANTS[name]
    return v[6] if len(v) > 6 else KINDS


def query_subset(name):
    """The queries of a variant (also used by the tests)."""
    qf, keep = VARIANTS[name][1], VARIANTS[name][3]
    return [r for i, r in enumerate(read_fasta(os.path.join(OUT, qf))) if keep(i)]


def write_inputs(name, tmp):
    df = VARIANTS[name][2]
    q, d = os.path.join(tmp, "q.fa"), os.path.join(tmp, "db.fa")
    with open(q, "w") as f:
        for lab, s in query_subset(name):
            f.write("%s\n%s\n" % (lab, s))
    with gzip.open(os.path.join(OUT, df), "rt") as fi, open(d, "w") as fo:
        fo.write(fi.read())
    return q, d


def write_fastq(path):
    """FASTQ form of fmt_nt's queries: deterministic qualities, CR LF line ends in every fifth record,
    blank lines at the end of the file (fastqseqsource.cpp:28-41 allows them only there)."""
    with open(path, "wb") as f:
        for i, (lab, s) in enumerate(query_subset("fmt_nt")):
            qual = "".join(chr(33 + (i * 7 + j * 13) % 41) for j in range(len(s)))
            eol = "\r\n" if i % 5 == 0 else "\n"
            f.write(("@%s%s%s%s+%s%s%s%s" % (lab[1:], eol, s, eol, lab[1:] if i % 2 else "", eol, qual, eol)).encode())
        f.write(b"\n\n")


def main_fastq():
    """fmt_fq: the fmt_nt command on FASTQ queries, plus -matchedfq / -notmatchedfq."""
    opts = VARIANTS["fmt_nt"][4]
    with tempfile.TemporaryDirectory() as tmp:
        _, d = write_inputs("fmt_nt", tmp)
        q = os.path.join(tmp, "q.fq")
        write_fastq(q)
        outs = {k: os.path.join(tmp, "o." + k) for k in ("hits", "matchedfq", "notmatchedfq", "matched", "uc")}
        subprocess.run([REF, "-usearch_global", q, "-db", d, "-threads", "1", "-quiet"] + opts + [
            "-userout", outs["hits"], "-userfields", HITFIELDS, "-matchedfq", outs["matchedfq"], "-notmatchedfq",
            outs["notmatchedfq"], "-matched", outs["matched"], "-uc", outs["uc"]], check=True, stdout=subprocess.DEVNULL,
            stderr=subprocess.DEVNULL)
        for k, path in outs.items():
            data = open(path, "rb").read()
            if k in ("hits", "matched", "uc"):  # the same queries: nothing new to store
                with gzip.open(os.path.join(OUT, "fmt_nt.%s.gz" % k), "rb") as f:
                    assert f.read() == data, "FASTQ and FASTA queries gave different " + k
                continue
            with gzip.GzipFile(os.path.join(OUT, "fmt_fq.%s.gz" % k), "wb", compresslevel=9, mtime=0) as f:
                f.write(data)
            print("golden fmt_fq", k, data.count(b"\n"), "lines")


def main():
    sums = {}
    for name in VARIANTS:
        cmd, opts, fields = VARIANTS[name][0], VARIANTS[name][4], VARIANTS[name][5]
        with tempfile.TemporaryDirectory() as tmp:
            q, d = write_inputs(name, tmp)
            base = [REF, "-" + cmd, q, "-db", d, "-threads", "1", "-quiet"] + opts
            outs = {k: os.path.join(tmp, "o." + k) for k in KINDS + ("dbm", "dbnm", "dbcut", "trim")}
            run = base + ["-userout", outs["user"], "-userfields", fields]
            for k, flag in FLAGS.items():
                run += [flag, outs[k]]
            subprocess.run(run, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            # (the hit table has fields that -output_no_hits refuses, userout.cpp:117-120)
            subprocess.run([x for x in base if x != "-output_no_hits"] + ["-userout", outs["hits"], "-userfields", HITFIELDS], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for k in kinds_of(name):
                data = open(outs[k], "rb").read()
                if k == "aln":  # the first two lines are the command line and the program/host line
                    data = b"\n".join(data.split(b"\n")[2:])
                with gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (name, k)), "wb", compresslevel=9, mtime=0) as f:
                    f.write(data)
                print("golden", name, k, data.count(b"\n"), "lines")
            for k in ("dbm", "dbnm", "dbcut") if len(VARIANTS[name]) == 6 else ():  # database files: kept as digests
                data = open(outs[k], "rb").read()
                sums["%s.%s" % (name, k)] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data),
                                             "seqs": data.count(b">")}
    with open(os.path.join(OUT, "fmt_db_sha256.json"), "w") as f:
        json.dump(sums, f, indent=1, sort_keys=True)
        f.write("\n")
    main_fastq()


if __name__ == "__main__":
    sys.exit(main())
