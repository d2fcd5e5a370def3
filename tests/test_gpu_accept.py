"""Accepter / Terminator / HitMgr options of -usearch_global through the host CLI: output files
byte-identical to the reference binary's (tools/make_golden_accept.py), one variant per rule group
(accepter.cpp:41-94,145-197; terminator.cpp:66-86; hitmgr.cpp:367-398; outputsink.cpp:392)."""
import gzip
import os
import subprocess
import sys

import pytest

from tests import util

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(util.ROOT, "tools"))


def _variants():
    import make_golden_accept
    return make_golden_accept.VARIANTS


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("acc")
    for name in ("db", "q"):
        with gzip.open(os.path.join(util.GOLDEN, "acc_%s.fa.gz" % name), "rb") as fi, open(d / (name + ".fa"), "wb") as fo:
            fo.write(fi.read())
    return d


@pytest.mark.parametrize("variant", ["acc_self", "acc_notself", "acc_selfid", "acc_cov", "acc_maxqcov", "acc_cols", "acc_size",
                                     "acc_qt", "acc_termid", "acc_termidd", "acc_maxhits", "acc_tophit", "acc_tophits",
                                     "acc_nohits"])
def test_accept_option_matches_reference_golden(inputs, variant):
    from usearch12_b200 import build
    cli = build.build_cli()
    out = {k: str(inputs / (variant + "." + k)) for k in ("user", "uc", "b6")}
    cmd = [cli, "-usearch_global", str(inputs / "q.fa"), "-db", str(inputs / "db.fa"), "-quiet", "-uc", out["uc"], "-blast6out",
           out["b6"], "-userout", out["user"], "-userfields", "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"]
    r = subprocess.run(cmd + _variants()[variant], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind in ("user", "uc", "b6"):
        got = open(out[kind]).read().splitlines()
        d = util.first_diff(got, util.golden_lines(variant, kind))
        assert d is None, (variant, kind, d)


def _variants2():
    import make_golden_accept
    return make_golden_accept.VARIANTS2


@pytest.mark.parametrize("variant", ["accl_nt_cov", "accl_nt_skew", "accl_aa_tcov", "accl_aa_maxid", "accg_aa_diffs", "accg_aa_qt"])
def test_accept_options_on_local_and_amino_acid_searches(variant, tmp_path):
    """The same Accepter rules in the other candidate loops (k_local, k_align<AA>): -usearch_local on
    nucleotides and proteins, amino acid -usearch_global; output files identical to the reference binary's."""
    from usearch12_b200 import build
    cli = build.build_cli()
    cmdname, qf, df, extra = _variants2()[variant]
    for src, dst in ((qf, "q.fa"), (df, "db.fa")):
        with gzip.open(os.path.join(util.GOLDEN, src), "rb") as fi, open(tmp_path / dst, "wb") as fo:
            fo.write(fi.read())
    out = {k: str(tmp_path / (variant + "." + k)) for k in ("user", "uc", "b6")}
    cmd = [cli, cmdname, str(tmp_path / "q.fa"), "-db", str(tmp_path / "db.fa"), "-quiet", "-uc", out["uc"], "-blast6out",
           out["b6"], "-userout", out["user"], "-userfields", "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand"]
    r = subprocess.run(cmd + extra, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind in ("user", "uc", "b6"):
        got = open(out[kind]).read().splitlines()
        d = util.first_diff(got, util.golden_lines(variant, kind))
        assert d is None, (variant, kind, d)


def test_local_refuses_the_rules_the_reference_crashes_on(tmp_path):
    from usearch12_b200 import build
    for src, dst in (("loc_nt_q.fa.gz", "q.fa"), ("loc_nt_db.fa.gz", "db.fa")):
        with gzip.open(os.path.join(util.GOLDEN, src), "rb") as fi, open(tmp_path / dst, "wb") as fo:
            fo.write(fi.read())
    r = subprocess.run([build.build_cli(), "-usearch_local", str(tmp_path / "q.fa"), "-db", str(tmp_path / "db.fa"), "-id", "0.8",
                        "-evalue", "1e-3", "-strand", "plus", "-minsl", "0.5", "-uc", str(tmp_path / "o.uc")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0 and "not supported with -usearch_local" in r.stdout


def test_accept_options_need_attributes():
    """-self without label identities fails loudly through the C ABI."""
    from usearch12_b200 import capi
    g = util.Golden()
    p = capi.default_params(id=0.9, self=True)
    s = capi.Searcher(capi.Index(g.db[:50], p, device=0), p)
    with pytest.raises(capi.UsbError):
        s.search(g.q[:10])
