""".udb databases on the GPU path: a database written by
-makeudb_usearch (the file is byte-identical to the reference's, tests/test_udb_cpu.py) searched
with -db x.udb must give the reference binary's output files, and the index built from its stored,
already masked sequences must hold exactly the rows of the file."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from tests import util
from usearch12_b200 import build, capi

pytestmark = pytest.mark.gpu


def _gunzip(name, dst):
    with gzip.open(os.path.join(util.GOLDEN, name), "rb") as fi, open(dst, "wb") as fo:
        fo.write(fi.read())
    return dst


def test_cli_search_against_udb_file_equals_reference(golden, tmp_path):
    cli = build.build_cli()
    q = _gunzip("q.fa.gz", str(tmp_path / "q.fa"))
    db = _gunzip("db.fa.gz", str(tmp_path / "db.fa"))
    udb = str(tmp_path / "db.udb")
    r = subprocess.run([cli, "-makeudb_usearch", db, "-output", udb], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    outs = {k: str(tmp_path / ("o." + k)) for k in ("user", "uc", "b6")}
    cmd = [cli, "-usearch_global", q, "-db", udb, "-quiet", "-id", "0.97", "-strand", "plus", "-userfields",
           "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand", "-userout", outs["user"], "-uc", outs["uc"],
           "-blast6out", outs["b6"]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind in ("user", "uc", "b6"):
        d = util.first_diff(open(outs[kind]).read().splitlines(), golden.lines("plus97", kind))
        assert d is None, "%s\n%s" % (kind, d)


def test_local_cli_search_against_amino_udb_equals_reference(tmp_path):
    cli = build.build_cli()
    g = util.GoldenLocal("aa")
    kw = dict(util.LOCAL_VARIANTS["loc_aa_e5"])
    q = _gunzip("loc_aa_q.fa.gz", str(tmp_path / "q.fa"))
    db = _gunzip("loc_aa_db.fa.gz", str(tmp_path / "db.fa"))
    udb = str(tmp_path / "db.udb")
    r = subprocess.run([cli, "-makeudb_usearch", db, "-output", udb], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = str(tmp_path / "o.b6")
    cmd = [cli, "-usearch_local", q, "-db", udb, "-quiet", "-id", str(kw["id"]), "-evalue", str(kw["evalue"]), "-blast6out", out]
    if "maxaccepts" in kw:
        cmd += ["-maxaccepts", str(kw["maxaccepts"]), "-maxrejects", str(kw["maxrejects"])]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    d = util.first_diff(open(out).read().splitlines(), g.lines("loc_aa_e5", "b6"))
    assert d is None, d


def test_index_from_udb_sequences_has_the_rows_of_the_file(golden, tmp_path):
    path = str(tmp_path / "db.udb")
    capi.udb_write(path, golden.db_labels, golden.db)
    u = capi.Udb(path)
    p = capi.default_params(dbmask=0)  # stored sequences are masked already (loaddb.cpp:107-118)
    ix = capi.Index(u.seqs, p, device=0)
    assert ix.posting_width == 2  # static DB: 2-byte increment descriptors on the device, ascending rows on the host
    rng = np.random.default_rng(11)
    for w in rng.integers(0, 65536, 400):
        assert np.array_equal(ix.row(int(w)), u.row(int(w))), int(w)
    for t in (0, 7, len(u.seqs) - 1):
        assert ix.seq(t) == u.seqs[t]
    ix.close()
    u.close()


def test_device_index_build_equals_host_build():
    """k_ix_pass / k_ix_scan (usb_ixbuild.inc) against the host builder (usb_hostindex.cpp): every row
    of a 5 000-target database with masked runs, wildcards, lower case and short targets."""
    import random
    rng = random.Random(17)
    roots = ["".join(rng.choice("ACGT") for _ in range(400)) for _ in range(40)]
    db = []
    for i in range(5000):
        s = util.mutate(roots[i % 40], rng.uniform(0.02, 0.2), rng)
        k = rng.random()
        if k < 0.05:
            s = s[:150] + "A" * 12 + "ACACACACACAC" + s[150:]      # masking triggers
        elif k < 0.10:
            s = s[:60] + "N" + s[61:200].lower() + "RY" + s[202:]  # wildcards, lower case
        elif k < 0.12:
            s = s[:rng.randrange(0, 12)]                            # shorter than a word / empty-ish
        db.append(s if s else "A")
    os.environ["USB_HOST_INDEX"] = "1"
    try:
        host = capi.Index(db, capi.default_params(), device=0)
    finally:
        del os.environ["USB_HOST_INDEX"]
    dev = capi.Index(db, capi.default_params(), device=0)
    assert host.posting_count == dev.posting_count and dev.posting_count > 100000
    for w in range(65536):
        a, b = host.row(w), dev.row(w)
        assert np.array_equal(a, b), w
    host.close()
    dev.close()
