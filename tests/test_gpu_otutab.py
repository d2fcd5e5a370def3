"""-otutab (a caller that reuses the search path as a library;
searchcmd.cpp:21-40, otutabsink.cpp, otutab.cpp:247-310, label.cpp:152-234).  Golden files come
from the reference binary with -threads 1 (tools/make_golden_otutab.py): reads and OTUs carry
size=, sample=, barcodelabel= and otu= annotations in every form the label parser understands."""
import gzip
import os
import subprocess

import pytest

from tests import util
from usearch12_b200 import build

pytestmark = pytest.mark.gpu


def _gunzip(name, dst):
    with gzip.open(os.path.join(util.GOLDEN, name), "rb") as fi, open(dst, "wb") as fo:
        fo.write(fi.read())
    return dst


def _golden(name):
    with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
        return f.read().splitlines()


@pytest.mark.parametrize("db_opt", ["-otus", "-db"])
def test_otutab_files_byte_identical_to_reference(db_opt, tmp_path):
    cli = build.build_cli()
    reads = _gunzip("otutab_reads.fa.gz", str(tmp_path / "reads.fa"))
    otus = _gunzip("otutab_otus.fa.gz", str(tmp_path / "otus.fa"))
    tab, mp = str(tmp_path / "tab.txt"), str(tmp_path / "map.txt")
    r = subprocess.run([cli, "-otutab", reads, db_opt, otus, "-otutabout", tab, "-mapout", mp, "-batch", "700"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "mapped to OTUs" in r.stdout
    d = util.first_diff(open(tab).read().splitlines(), _golden("otutab.tab.gz"))
    assert d is None, d
    d = util.first_diff(open(mp).read().splitlines(), _golden("otutab.map.gz"))
    assert d is None, d


def test_otutab_needs_an_otu_database(tmp_path):
    cli = build.build_cli()
    r = subprocess.run([cli, "-otutab", "x.fa", "-otutabout", "t.txt"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Must specify OTU FASTA -db, -otus or -zotus" in r.stdout
