"""GPU parity tests of the whole hot path (usb_search_batch) against the golden files written by
the unmodified reference binary, and against the oracle on fresh seeded inputs."""
import random

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", list(util.VARIANTS))
def test_search_matches_reference_golden(golden, variant):
    from usearch12_b200 import capi
    p = capi.default_params(**util.VARIANTS[variant])
    ix = capi.Index(golden.db, p)
    s = capi.Searcher(ix, p)
    res = s.search(golden.q)
    user, uc, b6 = util.product_lines(res, golden.q_labels, golden.q, golden.db_labels)
    for got, kind in ((user, "user"), (uc, "uc"), (b6, "b6")):
        d = util.first_diff(got, golden.lines(variant, kind))
        assert d is None, "%s %s\n%s" % (variant, kind, d)
    assert s.launch_count >= 2


def test_search_matches_oracle_seeded():
    """Fresh seeded DB/reads (not the golden inputs): product vs oracle, hit for hit."""
    from oracle import uso_py as O
    from usearch12_b200 import capi
    import sys, os
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    from gen_synth import generate
    db, reads = generate(ndb=1500, dblen=1200, nq=3000, qlen=250, seed=77, nroot=15)
    qlab = [r[0][1:] for r in reads]
    qs = [r[1] for r in reads]
    dlab = ["db%d" % i for i in range(len(db))]
    p = capi.default_params()
    ix = capi.Index(db, p)
    s = capi.Searcher(ix, p)
    res = s.search(qs)
    got = util.product_lines(res, qlab, qs, dlab)
    op = O.default_params()
    osr = O.Searcher(O.DB(db, op, dlab), op)
    want = util.oracle_lines(osr, qlab, qs, dlab)
    for a, b, kind in zip(got, want, ("user", "uc", "b6")):
        assert util.first_diff(a, b) is None, kind
    # search counters agree with what the sequential reference loop does
    assert int(res.qstat["n_accept"].sum()) == len(got[0])


def test_empty_and_degenerate_batches():
    from usearch12_b200 import capi
    rng = random.Random(1)
    db = ["".join(rng.choice("ACGT") for _ in range(300)) for _ in range(50)]
    p = capi.default_params()
    ix = capi.Index(db, p)
    s = capi.Searcher(ix, p)
    res = s.search([])
    assert len(res.hits) == 0
    res = s.search(["A", "ACGTACG", "N" * 50, db[3][20:220], "acgt" * 30])
    assert [int(res.qoff[i + 1] - res.qoff[i]) for i in range(5)] == [0, 0, 0, 1, 0]
    h = res.hits[0]
    assert int(h["target"]) == 3 and int(h["ids"]) == 200 and res.cigar(h) == "20I200M80I"


def test_idempotent_and_batch_split_invariant(golden):
    from usearch12_b200 import capi
    p = capi.default_params()
    ix = capi.Index(golden.db, p)
    s = capi.Searcher(ix, p)
    qs = golden.q[:600]
    a = s.search(qs)
    b = s.search(qs)
    assert np.array_equal(a.hits[["query", "target", "ids", "alnlen"]], b.hits[["query", "target", "ids", "alnlen"]])
    c1, c2 = s.search(qs[:250]), s.search(qs[250:])
    tgt = np.concatenate([c1.hits["target"], c2.hits["target"]])
    assert np.array_equal(a.hits["target"], tgt)


def _run_cli(tmp_path, golden, extra, outs):
    """Runs the C++ host driver (Searcher/HitMgr/OutputSink mirror) on the golden FASTA inputs."""
    import gzip
    import os
    import subprocess
    from usearch12_b200 import build
    cli = build.build_cli()
    paths = {}
    for name in ("q", "db"):
        dst = os.path.join(tmp_path, name + ".fa")
        with gzip.open(os.path.join(util.GOLDEN, name + ".fa.gz"), "rb") as fi, open(dst, "wb") as fo:
            fo.write(fi.read())
        paths[name] = dst
    cmd = [cli, "-usearch_global", paths["q"], "-db", paths["db"], "-quiet"] + extra
    for k in outs:
        paths[k] = os.path.join(tmp_path, "o." + k)
        cmd += [{"user": "-userout", "uc": "-uc", "b6": "-blast6out"}[k], paths[k]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return {k: open(paths[k]).read().splitlines() for k in outs}


@pytest.mark.parametrize("variant", ["plus97", "both80_ma0"])
def test_cli_output_files_byte_identical_to_reference(golden, variant, tmp_path):
    kw = util.VARIANTS[variant]
    extra = ["-id", str(kw["id"]), "-strand", "both" if kw["strand_both"] else "plus",
             "-userfields", "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand", "-batch", "1000"]
    if "maxaccepts" in kw:
        extra += ["-maxaccepts", str(kw["maxaccepts"]), "-maxrejects", str(kw["maxrejects"])]
    got = _run_cli(str(tmp_path), golden, extra, ["user", "uc", "b6"])
    for kind in ("user", "uc", "b6"):
        d = util.first_diff(got[kind], golden.lines(variant, kind))
        assert d is None, "%s %s\n%s" % (variant, kind, d)


def test_cli_extended_userfields_match_reference(golden, tmp_path):
    """tests/golden/both90x.user.gz: reference binary, -id 0.9 -strand both -maxaccepts 2
    -maxrejects 16 with 25 userfields (first/last-M coordinates, gap counts, full path ...)."""
    xf = ("query+target+id+fractid+dist+pairs+gaps+allgaps+qlot+qhit+qunt+tlot+thit+tunt+ql+tl+alnlen+opens+exts+"
          "aln+tstrand+mism+ids+diffs+clusternr")
    got = _run_cli(str(tmp_path), golden, ["-id", "0.9", "-strand", "both", "-maxaccepts", "2", "-maxrejects", "16",
                                           "-userfields", xf], ["user"])
    d = util.first_diff(got["user"], golden.lines("both90x", "user"))
    assert d is None, d


def test_cli_refuses_unsupported_options(tmp_path):
    import subprocess
    from usearch12_b200 import build
    cli = build.build_cli()
    r = subprocess.run([cli, "-usearch_global", "x.fa", "-db", "y.fa", "-id", "0.9", "-strand", "plus", "-uparse_break", "3"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "not supported" in r.stdout


def test_index_append_equals_one_shot_build(golden):
    """A DB grown by usb_index_append (log-structured CSR segments) searches exactly like the same
    DB built in one call; rows seen through usb_index_row are identical too."""
    from usearch12_b200 import capi
    p = capi.default_params()
    one = capi.Index(golden.db, p)
    grown = capi.Index(golden.db[:5], p)
    cuts = [5, 6, 40, 41, 200, 390, len(golden.db)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        grown.append(golden.db[a:b])
    rng = random.Random(2)
    for w in [rng.randrange(65536) for _ in range(500)]:
        assert np.array_equal(one.row(w), grown.row(w)), w
    qs = golden.q[:300] + golden.q[2400:2500]
    r1 = capi.Searcher(one, p).search(qs)
    r2 = capi.Searcher(grown, p).search(qs)
    assert np.array_equal(r1.hits[["query", "target", "ids", "alnlen", "rank"]], r2.hits[["query", "target", "ids", "alnlen", "rank"]])
    assert [r1.cigar(h) for h in r1.hits] == [r2.cigar(h) for h in r2.hits]


@pytest.mark.parametrize("mode", ["band0", "fulldp"])
def test_full_dp_variants_match_reference_golden(golden, mode):
    """-band 0 (ViterbiFastMem for the holes, globalalignmem.cpp:103-106) and -fulldp (no HSPs,
    full Viterbi per candidate, :153-157): reference-binary outputs on 800 golden reads."""
    from usearch12_b200 import capi
    kw = dict(band=0) if mode == "band0" else dict(fulldp=1)
    p = capi.default_params(**kw)
    ix = capi.Index(golden.db, p)
    s = capi.Searcher(ix, p)
    qs = golden.q[:600] + golden.q[-200:]
    labels = golden.q_labels[:600] + golden.q_labels[-200:]
    res = s.search(qs)
    user, uc, _ = util.product_lines(res, labels, qs, golden.db_labels)
    for got, kind in ((user, "user"), (uc, "uc")):
        d = util.first_diff(got, golden.lines(mode + "_q600", kind))
        assert d is None, "%s %s\n%s" % (mode, kind, d)
