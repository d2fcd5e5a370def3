#!/usr/bin/env python3
"""Golden fixtures for the alignment score options of -usearch_global (tests/golden/sc_*): -match, -mismatch
(alnparams.cpp:330-334), -minhsp, -xdrop_nw, -hspw, -band (alnheuristics.cpp:26-61), -bump, written by the UNMODIFIED reference
binary (oracle/_ref/usearch12) on the queries of the fmt_nt fixture.   Usage: python tools/make_golden_scores.py"""
import gzip
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(HERE, "tools"))
import make_golden_formats as M  # noqa: E402

USERFIELDS = "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"
BASE = ["-id", "0.9", "-strand", "both", "-maxaccepts", "2", "-maxrejects", "16"]
# name -> (command line options, the same as usb_params / oracle fields)
VARIANTS = {
    "sc_a": (["-match", "2", "-mismatch", "-3", "-minhsp", "24", "-xdrop_nw", "12"],
             dict(match=2.0, mismatch=-3.0, minhsp=24, xdrop_nw=12.0)),
    "sc_b": (["-match", "1", "-mismatch", "-1", "-minhsp", "12", "-xdrop_nw", "4"],
             dict(match=1.0, mismatch=-1.0, minhsp=12, xdrop_nw=4.0)),
    "sc_c": (["-match", "3", "-mismatch", "-1", "-xdrop_nw", "20"], dict(match=3.0, mismatch=-1.0, xdrop_nw=20.0)),
    # HSP word length (alnheuristics.cpp:60-61), band radius (:33), U-sort bump (udbusortedsearcher.cpp:269-282)
    "sc_d": (["-hspw", "4", "-band", "8", "-bump", "0"], dict(hspw=4, band=8, bump=0)),
    "sc_e": (["-hspw", "6", "-band", "40", "-bump", "80"], dict(hspw=6, band=40, bump=80)),
}
PARAMS = dict(id=0.9, strand_both=1, maxaccepts=2, maxrejects=16)


def main():
    for name, (opts, _) in VARIANTS.items():
        with tempfile.TemporaryDirectory() as tmp:
            q, d = M.write_inputs("fmt_nt", tmp)
            outs = {k: os.path.join(tmp, "o." + k) for k in ("user", "uc", "b6")}
            subprocess.run([M.REF, "-usearch_global", q, "-db", d, "-threads", "1", "-quiet"] + BASE + opts + [
                "-userout", outs["user"], "-userfields", USERFIELDS, "-uc", outs["uc"], "-blast6out", outs["b6"]],
                check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for k, path in outs.items():
                data = open(path, "rb").read()
                with gzip.GzipFile(os.path.join(M.OUT, "%s.%s.gz" % (name, k)), "wb", compresslevel=9, mtime=0) as f:
                    f.write(data)
                print("golden", name, k, data.count(b"\n"), "lines")


if __name__ == "__main__":
    sys.exit(main())
