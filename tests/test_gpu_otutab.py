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


def test_closed_ref_sink_follows_the_reference_source(tmp_path):
    """-closed_ref (searchcmd.cpp:11-19, closedrefsink.cpp:34-165).  PARITY UNPINNED: both reference
    binaries (the one built here and the prebuilt tmp/usearch_linux_x86_12.0-beta) crash with SIGSEGV on
    -closed_ref, so there are no golden files; the sink is checked against a restatement of
    ClosedRefSink::OnQueryDone / OnAllDone applied to the hit lists of the same run (-userout)."""
    import numpy as np
    cli = build.build_cli()
    db = _gunzip("acc_db.fa.gz", str(tmp_path / "db.fa"))
    q = _gunzip("acc_q.fa.gz", str(tmp_path / "q.fa"))
    out = {k: str(tmp_path / k) for k in ("tab", "dbotus", "dataotus", "user")}
    r = subprocess.run([cli, "-closed_ref", q, "-db", db, "-strand", "both", "-id", "0.9", "-quiet", "-tabbedout", out["tab"],
                        "-dbotus", out["dbotus"], "-dataotus", out["dataotus"], "-userout", out["user"], "-userfields",
                        "query+target+ids+alnlen+clusternr"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    ql, qs = util.read_fasta(q)
    # (labels occur twice among the queries: -userout lines of one query are consecutive, a query
    # without hits writes none -- cut the file into runs and give them to the queries in order)
    runs = []
    for line in open(out["user"]).read().splitlines():      # HitMgr order per query
        qq, t, ids, aln, ti = line.split("\t")
        if not runs or runs[-1][0] != qq:
            runs.append((qq, []))
        runs[-1][1].append((t, np.float32(int(ids) / int(aln)), int(ti)))
    dup = {x for x in ql if ql.count(x) > 1}
    hits = {qq: h for qq, h in runs if qq not in dup}
    want, otu_of, totals, members = [], {}, [], []
    got = open(out["tab"]).read().splitlines()
    assert len(got) == len(ql)
    for n, qq in enumerate(ql):
        if qq in dup:                                          # take the product's line: the OTU bookkeeping must go on
            f = got[n].split("\t")
            want.append(got[n])
            if f[1] != "*":
                ti = [k for k, lab in enumerate(util.read_fasta(db)[0]) if lab == f[3]][0]
                if ti not in otu_of:
                    otu_of[ti] = len(totals)
                    totals.append(0)
                    members.append(0)
                totals[otu_of[ti]] += int(qq.split(";size=")[1].split(";")[0]) if ";size=" in qq else 1
                members[otu_of[ti]] += 1
            continue
        h = hits.get(qq)
        if not h:
            want.append("%s\t*\t*\t*\t*\t*" % qq)
            continue
        top = h[0]
        for x in h:                                            # GetTopHit: best score, ties to the lowest target index
            if x[1] > top[1] or (x[1] == top[1] and x[2] < top[2]):
                top = x
        if top[2] not in otu_of:
            otu_of[top[2]] = len(totals)
            totals.append(0)
            members.append(0)
        o = otu_of[top[2]]
        totals[o] += int(qq.split(";size=")[1].split(";")[0]) if ";size=" in qq else 1
        m = members[o]
        members[o] += 1
        ties = []
        if len(h) > 1:
            for x in h:
                if x[1] < h[0][1]:
                    break
                if x[2] != top[2]:
                    ties.append(x[0])
        line = "%s\t%d\t%d\t%s\t%.1f\tties=%d" % (qq, o, m, top[0], float(h[0][1]) * 100.0, len(ties))
        want.append(line + (":" + ",".join(ties) if ties else ""))
    assert util.first_diff(got, want) is None
    assert len(totals) > 50
    # OTU files: every OTU once, labels ...otu=<rank>;size=<total>; in decreasing size order
    dl, _ = util.read_fasta(out["dbotus"])
    al, _ = util.read_fasta(out["dataotus"])
    assert len(dl) == len(totals) == len(al)
    sizes = [int(x.rstrip(";").split(";size=")[-1]) for x in dl]
    assert sizes == sorted(sizes, reverse=True) and sorted(sizes) == sorted(totals)
    for k, (d, a_) in enumerate(zip(dl, al)):
        assert ";otu=%d;" % (k + 1) in d and ";otu=%d;ref=" % (k + 1) in a_
