"""GPU parity tests of the -usearch_local path (config 5): amino acid / nucleotide UDB ranking,
LocalAligner2 seeding, X-drop gapped extension and E-value gates through the C ABI, against the
golden files written by the unmodified reference binary (tools/make_golden_local.py) and against
the oracle on fresh seeded inputs."""
import os
import random
import sys

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(util.ROOT, "tools"))


def _searchers(db, nucleo, evalue, **kw):
    from usearch12_b200 import capi
    p = util.product_local_params(nucleo, evalue, **kw)
    ix = capi.Index(db, p)
    return ix, capi.Searcher(ix, p)


@pytest.mark.parametrize("variant", list(util.LOCAL_VARIANTS))
def test_local_search_matches_reference_golden(variant):
    kw = dict(util.LOCAL_VARIANTS[variant])
    nucleo = kw.pop("nucleo")
    evalue = kw.pop("evalue")
    g = util.GoldenLocal("nt" if nucleo else "aa")
    ix, s = _searchers(g.db, nucleo, evalue, **kw)
    res = s.search(g.q)
    got = util.product_lines_local(res, s, g.q_labels, g.q, g.db_labels, nucleo)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        d = util.first_diff(lines, g.lines(variant, kind))
        assert d is None, "%s %s\n%s" % (variant, kind, d)
    assert s.launch_count >= 2


def test_amino_index_and_rank_match_oracle():
    """a1-a6 on the 20-letter alphabet: UDB rows (5-mers, 3.2 M slots), U vectors, candidate order."""
    from oracle import uso_py as O
    import gen_synth_aa
    db, qs = gen_synth_aa.generate(ndb=700, length=300, nq=60, seed=21, nroot=10)
    db[3] = db[3][:50] + "A" * 12 + db[3][62:]
    db[5] = db[5][:120].lower() + db[5][120:]
    db[7] = db[7][:30] + "XBZ" + db[7][33:]
    queries = [q for _, q in qs] + ["ACDEF", "ACDE", "", "X" * 50, db[9][:90].lower() + db[9][90:], db[11] * 2]
    ix, s = _searchers(db, False, 1e-5, id=0.5)
    op = util.oracle_local_params(False, id=0.5, evalue=1e-5)
    odb = O.DB(db, op)
    osr = O.Searcher(odb, op)
    for t in (0, 3, 5, 7, 699):
        assert ix.seq(t) == odb.seq(t)
    rng = random.Random(3)
    words = [rng.randrange(20 ** 5) for _ in range(300)]
    # words that certainly occur
    aa = "ACDEFGHIKLMNPQRSTVWY"
    for t in (0, 100, 650):
        for pos in (0, 17, 200):
            w = 0
            for c in db[t][pos:pos + 5]:
                w = w * 20 + aa.index(c)
            words.append(w)
    for w in words:
        assert np.array_equal(ix.row(w), odb.row(w)), w
    k_max = 64
    ct, cu, nc, u = s.rank(queries, k_max, want_u=True)
    for i, q in enumerate(queries):
        U, ot, ou = osr.rank(q)
        assert np.array_equal(u[i], U), i
        assert nc[i] == len(ot), i
        n = min(k_max, len(ot))
        assert np.array_equal(ct[i, :n], ot[:n]) and np.array_equal(cu[i, :n], ou[:n]), i


@pytest.mark.parametrize("nucleo", [False, True])
def test_local_search_matches_oracle_seeded(nucleo):
    """Fresh seeded inputs (not the golden ones): product vs oracle, hit for hit."""
    from oracle import uso_py as O
    if nucleo:
        import gen_synth
        db, reads = gen_synth.generate(ndb=800, dblen=1000, nq=1500, qlen=300, seed=91, nroot=10)
        qlab = [r[0][1:] for r in reads]
        qs = [r[1] for r in reads]
        dlab = ["db%d" % i for i in range(len(db))]
        kw = dict(id=0.8, maxaccepts=3, maxrejects=16)
        evalue = 1e-3
    else:
        import gen_synth_aa
        db, recs = gen_synth_aa.generate(ndb=1500, length=350, nq=2500, seed=92, nroot=40)
        qlab = [r[0] for r in recs]
        qs = [r[1] for r in recs]
        dlab = ["p%d" % i for i in range(len(db))]
        kw = dict(id=0.4, maxaccepts=2, maxrejects=32)
        evalue = 1e-4
    ix, s = _searchers(db, nucleo, evalue, **kw)
    res = s.search(qs)
    got = util.product_lines_local(res, s, qlab, qs, dlab, nucleo)
    op = util.oracle_local_params(nucleo, evalue=evalue, **kw)
    osr = O.Searcher(O.DB(db, op, dlab), op)
    want = util.oracle_lines_local(osr, qlab, qs, dlab, nucleo)
    for a, b, kind in zip(got, want, ("user", "uc", "b6")):
        assert util.first_diff(a, b) is None, kind
    assert len(got[0]) > len(qs) // 2


def test_local_degenerate_batches_and_split_invariance():
    import gen_synth_aa
    db, recs = gen_synth_aa.generate(ndb=300, length=200, nq=200, seed=5, nroot=10)
    qs = [r[1] for r in recs] + ["", "A", "ACD", "ACDEFG", "W" * 300]
    ix, s = _searchers(db, False, 1e-5, id=0.5)
    assert len(s.search([]).hits) == 0
    whole = s.search(qs)
    parts = [s.search(qs[:77]), s.search(qs[77:])]
    key = lambda r: [(int(h["target"]), int(h["raw"]), int(h["first_mq"]), int(h["last_mt"]), r.cigar(h)) for h in r.hits]
    assert key(whole) == key(parts[0]) + key(parts[1])
    again = s.search(qs)
    assert key(whole) == key(again)
    # sequences above g_MaxL letters run in the long mode (split extensions); beyond the 16-bit position
    # limit they are refused loudly, not silently mis-aligned
    from usearch12_b200 import capi
    assert len(s.search(["ACDEFGHIKL" * 500]).qoff) == 2
    with pytest.raises(capi.UsbError):
        s.search(["ACDEFGHIKL" * 7000])


def test_local_pairs_stage():
    """usb_local_pairs: every AR of explicit pairs; a circularly permuted query gives two ARs."""
    import gen_synth_aa
    rng = random.Random(8)
    db, _ = gen_synth_aa.generate(ndb=50, length=300, nq=1, seed=6, nroot=50)
    qs = [db[3][150:] + db[3][:150], db[7], gen_synth_aa.mutate(db[9], 0.2, rng), "".join(rng.choice(gen_synth_aa.AA) for _ in range(200))]
    ix, s = _searchers(db, False, 10.0, id=0.0)
    res = s.local_pairs(qs, [0, 1, 2, 3, 1], [3, 7, 9, 11, 8])
    n = [int(res.qoff[i + 1] - res.qoff[i]) for i in range(5)]
    assert n[0] == 2 and n[1] == 1 and n[2] == 1 and n[3] == 0 and n[4] == 0, n
    h = res.hits[int(res.qoff[1])]
    assert res.cigar(h) == "300M" and int(h["ids"]) == 300
    a, b = res.hits[0], res.hits[1]
    assert {(int(a["first_mq"]), int(a["first_mt"])), (int(b["first_mq"]), int(b["first_mt"]))} == {(0, 150), (150, 0)}


@pytest.mark.parametrize("variant", list(util.LOCAL_LONG_VARIANTS))
def test_local_split_extensions_match_reference_golden(variant):
    """Sequences above g_MaxL = 4096 letters (up to 20 000 here): extensions longer than g_MaxL are split
    (xdropalignmem.cpp:87-139 -> XDropFwdSplit / XDropBwdSplit); letters and word keys of the long mode
    live in the per-warp slab.  Output lines identical to the reference binary's."""
    kw = dict(util.LOCAL_LONG_VARIANTS[variant])
    nucleo = kw.pop("nucleo")
    evalue = kw.pop("evalue")
    g = util.GoldenLocal("nt" if nucleo else "aa", prefix="loclong")
    ix, s = _searchers(g.db, nucleo, evalue, **kw)
    res = s.search(g.q)
    got = util.product_lines_local(res, s, g.q_labels, g.q, g.db_labels, nucleo)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        d = util.first_diff(lines, g.lines(variant, kind))
        assert d is None, "%s %s\n%s" % (variant, kind, d)
    assert max(int(l.split("\t")[3]) for l in got[0]) > 4096


def test_local_long_and_short_sequences_in_one_database():
    """A batch whose longest sequence is above g_MaxL switches the whole batch to the long mode: the short
    pairs in it must give what the normal mode gives (oracle, seeded inputs)."""
    from oracle import uso_py as O
    rng = random.Random(31)
    g = util.GoldenLocal("nt")
    db = list(g.db[:60]) + ["".join(rng.choice("ACGT") for _ in range(9000))]
    qs = list(g.q[:80]) + [util.mutate(db[-1][1000:8000], 0.05, rng)]
    labels = ["q%d" % i for i in range(len(qs))]
    dbl = ["t%d" % i for i in range(len(db))]
    ix, s = _searchers(db, True, 1e-5, id=0.8, strand_both=1)
    res = s.search(qs)
    got = util.product_lines_local(res, s, labels, qs, dbl, True)
    op = util.oracle_local_params(True, id=0.8, evalue=1e-5, strand_both=1)
    osr = O.Searcher(O.DB(db, op, dbl), op)
    want = util.oracle_lines_local(osr, labels, qs, dbl, True)
    for a, b, kind in zip(got, want, ("user", "uc", "b6")):
        assert util.first_diff(a, b) is None, kind
    assert len(got[0]) > 40


@pytest.mark.parametrize("variant", ["loc_aa_e5", "loc_nt_both"])
def test_local_cli_output_files_byte_identical_to_reference(variant, tmp_path):
    """The C++ host driver (usearch12_b200_cli -usearch_local) writes the reference's files."""
    import gzip
    import subprocess
    from usearch12_b200 import build
    kw = dict(util.LOCAL_VARIANTS[variant])
    nucleo = kw.pop("nucleo")
    kind = "nt" if nucleo else "aa"
    g = util.GoldenLocal(kind)
    cli = build.build_cli()
    paths = {}
    for name in ("q", "db"):
        dst = os.path.join(str(tmp_path), name + ".fa")
        with gzip.open(os.path.join(util.GOLDEN, "loc_%s_%s.fa.gz" % (kind, name)), "rb") as fi, open(dst, "wb") as fo:
            fo.write(fi.read())
        paths[name] = dst
    cmd = [cli, "-usearch_local", paths["q"], "-db", paths["db"], "-quiet", "-id", str(kw["id"]), "-evalue", str(kw["evalue"]),
           "-userfields", "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand", "-batch", "300"]
    if nucleo:
        cmd += ["-strand", "both" if kw.get("strand_both") else "plus"]
    if "maxaccepts" in kw:
        cmd += ["-maxaccepts", str(kw["maxaccepts"]), "-maxrejects", str(kw["maxrejects"])]
    for k, o in (("user", "-userout"), ("uc", "-uc"), ("b6", "-blast6out")):
        paths[k] = os.path.join(str(tmp_path), "o." + k)
        cmd += [o, paths[k]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for k in ("user", "uc", "b6"):
        d = util.first_diff(open(paths[k]).read().splitlines(), g.lines(variant, k))
        assert d is None, "%s %s\n%s" % (variant, k, d)
