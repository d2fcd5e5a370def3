#!/usr/bin/env python3
"""Digests of the .udb files the UNMODIFIED reference binary (oracle/_ref/usearch12 -makeudb_usearch)
writes for the golden databases -> tests/golden/udb_sha256.json.  The files themselves (3 MB and
13 MB) are not committed; tests/test_udb_cpu.py compares the digest of usb_udb_write's output, and
the bytes themselves wherever the reference binary is present."""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "usearch12")
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name in ("db.fa.gz", "loc_aa_db.fa.gz", "loc_nt_db.fa.gz"):
        fa = os.path.join(tmp, name[:-3])
        with gzip.open(os.path.join(ROOT, "tests", "golden", name), "rb") as f, open(fa, "wb") as g:
            g.write(f.read())
        udb = fa + ".udb"
        subprocess.run([REF, "-makeudb_usearch", fa, "-output", udb, "-quiet"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        b = open(udb, "rb").read()
        out[name] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
    # -wordlength (udbparams.cpp:58-81): keys "<file>:w<length>"
    for name, w in (("db.fa.gz", 6), ("loc_aa_db.fa.gz", 4)):
        fa = os.path.join(tmp, name[:-3])
        udb = fa + ".w.udb"
        subprocess.run([REF, "-makeudb_usearch", fa, "-output", udb, "-wordlength", str(w), "-quiet"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        b = open(udb, "rb").read()
        out["%s:w%d" % (name, w)] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "udb_sha256.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
