#!/usr/bin/env python3
"""Golden fixtures for amino acid -usearch_global (BASELINE config 1 and variants):
outputs of the UNMODIFIED reference binary (oracle/_ref/usearch12).

  tests/golden/cfg1_test.fa.gz        the reference's own tmp/test.fa (266 SCOP domains; data fixture)
  tests/golden/cfg1_id90.*            config 1 as stated: test.fa vs itself, -id 0.9 -threads 1
  tests/golden/cfg1_id30_ma8.*        -id 0.3 -maxaccepts 8 -maxrejects 64
  tests/golden/gaa_*.{user,uc,b6}.gz  the protein families of tests/golden/loc_aa_{db,q}.fa.gz (wildcards
                                      B/Z/X/U/O/J, lower case, masked runs, chimeras, fragments, short
                                      queries) searched globally at four -id / Terminator settings

Only runs where the reference binary and /root/reference exist (the build container).
Usage: python tools/make_golden_aa_global.py
"""
import gzip
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(HERE, "oracle", "_ref", "usearch12")
OUT = os.path.join(HERE, "tests", "golden")
USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"

VARIANTS = {
    "gaa_id50": ["-id", "0.5"],
    "gaa_id30_ma8": ["-id", "0.3", "-maxaccepts", "8", "-maxrejects", "64"],
    "gaa_id90": ["-id", "0.9"],
    "gaa_id70_ma3": ["-id", "0.7", "-maxaccepts", "3", "-maxrejects", "16"],
}
CFG1 = {
    "cfg1_id90": ["-id", "0.9"],
    "cfg1_id30_ma8": ["-id", "0.3", "-maxaccepts", "8", "-maxrejects", "64"],
}


def gunzip_to(src, dst):
    with gzip.open(src, "rb") as fi, open(dst, "wb") as fo:
        fo.write(fi.read())


def run(name, q, db, extra, tmp):
    cmd = [REF, "-usearch_global", q, "-db", db, "-threads", "1", "-quiet", "-uc", os.path.join(tmp, "uc"), "-blast6out",
           os.path.join(tmp, "b6"), "-userout", os.path.join(tmp, "user"), "-userfields", USERFIELDS] + extra
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for x in ("user", "uc", "b6"):
        with open(os.path.join(tmp, x), "rb") as fi, gzip.GzipFile(os.path.join(OUT, "%s.%s.gz" % (name, x)), "wb",
                                                                    compresslevel=9, mtime=0) as fo:
            fo.write(fi.read())
    print("golden", name)


def main():
    tmp = tempfile.mkdtemp()
    try:
        src = "/root/reference/tmp/test.fa"
        with open(src, "rb") as fi, gzip.GzipFile(os.path.join(OUT, "cfg1_test.fa.gz"), "wb", compresslevel=9, mtime=0) as fo:
            fo.write(fi.read())
        for name, extra in CFG1.items():
            run(name, src, src, extra, tmp)
        db, q = os.path.join(tmp, "db.fa"), os.path.join(tmp, "q.fa")
        gunzip_to(os.path.join(OUT, "loc_aa_db.fa.gz"), db)
        gunzip_to(os.path.join(OUT, "loc_aa_q.fa.gz"), q)
        for name, extra in VARIANTS.items():
            run(name, q, db, extra, tmp)
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
