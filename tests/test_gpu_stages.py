"""GPU parity tests, stage by stage, through the C ABI (libusb200.so) against the CPU oracle."""
import random

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(golden):
    from oracle import uso_py as O
    from usearch12_b200 import capi
    p = capi.default_params()
    ix = capi.Index(golden.db, p)
    op = O.default_params()
    odb = O.DB(golden.db, op, golden.db_labels)
    return dict(capi=capi, O=O, ix=ix, odb=odb, g=golden)


def test_index_rows_and_masking(env):
    ix, odb, g = env["ix"], env["odb"], env["g"]
    for t in range(len(g.db)):
        assert ix.seq(t) == odb.seq(t), t
    rng = random.Random(5)
    words = [rng.randrange(65536) for _ in range(3000)] + [0, 65535, 0x1111, 0x4444]
    total = 0
    for w in words:
        a, b = ix.row(w), odb.row(w)
        assert np.array_equal(a, b), w
        total += len(a)
    assert total > 0


@pytest.mark.parametrize("strand_both", [0, 1])
def test_rank_matches_oracle(env, strand_both):
    capi, O, g = env["capi"], env["O"], env["g"]
    p = capi.default_params(strand_both=strand_both)
    s = capi.Searcher(env["ix"], p)
    osr = O.Searcher(env["odb"], O.default_params())
    idx = list(range(0, 200)) + list(range(2400, len(g.q)))
    seqs = [g.q[i] for i in idx]
    K = 40
    ct, cu, nc, U = s.rank(seqs, K, want_u=True)
    strands = 2 if strand_both else 1
    for n, q in enumerate(seqs):
        for st in range(strands):
            qq = q if st == 0 else O.revcomp(q)
            oU, oct, ocu = osr.rank(qq)
            j = n * strands + st
            assert np.array_equal(U[j], oU), (idx[n], st, "U")
            assert nc[j] == len(oct), (idx[n], st, nc[j], len(oct))
            k = min(K, len(oct))
            assert np.array_equal(ct[j, :k], oct[:k]), (idx[n], st, ct[j, :k], oct[:k])
            assert np.array_equal(cu[j, :k], ocu[:k]), (idx[n], st)


def test_rank_large_k_and_ties(env):
    """k_max larger than the candidate count and heavy ties (many identical targets)."""
    capi, O = env["capi"], env["O"]
    rng = random.Random(9)
    base = "".join(rng.choice("ACGT") for _ in range(400))
    db = [base[:300 + (i % 7)] for i in range(300)] + [util.mutate(base, 0.05, rng) for _ in range(300)]
    p = capi.default_params(dbmask=1)
    ix = capi.Index(db, p)
    s = capi.Searcher(ix, p)
    odb = O.DB(db, O.default_params())
    osr = O.Searcher(odb)
    qs = [base[10:260], util.mutate(base[50:300], 0.03, rng), base[100:130], "ACGT" * 20]
    for K in (8, 64, 1024):
        ct, cu, nc, U = s.rank(qs, K, want_u=True)
        for j, q in enumerate(qs):
            oU, oct, ocu = osr.rank(q)
            assert np.array_equal(U[j], oU)
            assert nc[j] == len(oct)
            k = min(K, len(oct))
            assert np.array_equal(ct[j, :k], oct[:k]), (K, j)


def test_viterbi_matches_oracle(env):
    capi, O = env["capi"], env["O"]
    s = capi.Searcher(env["ix"], env["ix"].params)
    rng = random.Random(3)
    A, B, F = [], [], []
    for k in range(640):
        la = rng.choice([1, 2, 3, 5, 17, 24, 31, 32, 33, 40, 64, 65, 100, 184, 250])
        # (rows of 96 and more columns take the four-columns-per-lane sweep)
        b = "".join(rng.choice("ACGT") for _ in range(rng.choice([1, 2, 7, 33, 64, 95, 96, 97, 100, 127, 128, 129, 131, 255, 257,
                                                                  364, 511, 700, 1291, 1409])))
        mode = k % 4
        if mode == 0:      # related: A is a mutated window of B
            st = rng.randrange(0, max(1, len(b) - la + 1))
            a = util.mutate(b[st:st + la], 0.1, rng) or "A"
        elif mode == 1:    # unrelated
            a = "".join(rng.choice("ACGT") for _ in range(la))
        elif mode == 2:    # B is a mutated version of A-like, similar lengths, with wildcards
            a = "".join(rng.choice("ACGT") for _ in range(la))
            b = util.mutate(a, 0.15, rng) or "C"
            if len(b) > 3:
                b = b[:1] + "N" + b[2:]
        else:              # A longer than B
            a = "".join(rng.choice("ACGT") for _ in range(la + len(b) % 50))
            b = b[:max(1, la // 2)]
        A.append(a); B.append(b); F.append(rng.randrange(16))
    paths, sc = s.viterbi(A, B, F)
    op = O.default_params()
    for k in range(len(A)):
        f = F[k]
        opath, osc = O.viterbi_band(op, A[k], B[k], f & 1, f & 2, f & 4, f & 8)
        assert paths[k] == opath, (k, len(A[k]), len(B[k]), f, paths[k][:80], opath[:80])
        assert sc[k] == int(round(2 * osc)), (k, sc[k], osc)


def test_align_pairs_match_oracle(env):
    capi, O, g = env["capi"], env["O"], env["g"]
    s = capi.Searcher(env["ix"], env["ix"].params)
    osr = O.Searcher(env["odb"], O.default_params())
    rng = random.Random(11)
    idx = list(range(0, 120)) + list(range(2400, len(g.q)))
    seqs = [g.q[i] for i in idx]
    pq, pt = [], []
    for n, i in enumerate(idx):
        lab = g.q_labels[i]
        tt = [rng.randrange(len(g.db))]
        if "t=db" in lab:
            t = int(lab.split("t=db")[1].split(";")[0])
            tt += [t, (t + 8) % 400]
        for t in tt:
            pq.append(n); pt.append(t)
    aligned, res, hsp = s.align_pairs(seqs, pq, pt, max_hsp=32)
    n_al = 0
    for k in range(len(pq)):
        q, t = seqs[pq[k]], env["odb"].seq(pt[k])
        ung, ch, fid = osr.global_hsps(q, t)
        opath = osr.global_align(q, t)
        nch = int(hsp[k, 0])
        assert nch == len(ch), (k, nch, len(ch))
        got = hsp[k, 1:1 + 4 * nch].reshape(-1, 4)
        assert np.array_equal(got, ch[:nch]), (k, got, ch)
        assert bool(aligned[k]) == (opath is not None), (k, aligned[k], opath is None)
        if opath is not None:
            b, e = int(res.qoff[k]), int(res.qoff[k + 1])
            assert e - b == 1
            assert res.path(res.hits[b]) == opath, (k, idx[pq[k]], pt[k])
            n_al += 1
    assert n_al > 50


@pytest.mark.parametrize("big,id_,stepwords", [(100, 0.97, 8), (100, 0.8, 8), (50, 0.97, 0)])
def test_big_database_path_matches_oracle(env, big, id_, stepwords):
    """UDBSearchBig (udbusortedsearcherbig.cpp): DB larger than -big; candidate order and whole
    searches against the oracle, which is pinned on this path by tools/pin_oracle.sh with a
    120 000-sequence DB."""
    capi, O, g = env["capi"], env["O"], env["g"]
    p = capi.default_params(big=big, id=id_, stepwords=stepwords)
    ix = capi.Index(g.db, p)
    s = capi.Searcher(ix, p)
    op = O.default_params(big=big, id=id_, stepwords=stepwords)
    odb = O.DB(g.db, op, g.db_labels)
    osr = O.Searcher(odb, op)
    idx = list(range(0, 150)) + list(range(2400, 2700, 3))
    seqs = [q for q in (g.q[i] for i in idx) if len(q) <= 4000]
    labels = ["q%d" % i for i in range(len(seqs))]
    res = s.search(seqs)
    got = util.product_lines(res, labels, seqs, g.db_labels)
    want = util.oracle_lines(osr, labels, seqs, g.db_labels)
    for a, b, kind in zip(got, want, ("user", "uc", "b6")):
        assert util.first_diff(a, b) is None, kind
    assert len(got[0]) > 50


@pytest.mark.parametrize("stepwords", [0, 8])
def test_big_path_with_more_survivors_than_slots(env, stepwords):
    """UDBSearchBig on a database where thousands of targets pass U >= NextValue/2 (queries without a
    close target: NextValue is 1 or 2): the first k_max of (U descending, first-touch order) are
    selected by the histogram of the survivor pass, ties at the cut row by row
    (udbusortedsearcherbig.cpp:82-135, countsort.cpp:110-191)."""
    capi, O = env["capi"], env["O"]
    rng = random.Random(77 + stepwords)
    db = ["".join(rng.choice("ACGT") for _ in range(300)) for _ in range(60000)]
    p = capi.default_params(big=1000, stepwords=stepwords)
    ix = capi.Index(db, p)
    s = capi.Searcher(ix, p)
    op = O.default_params(big=1000, stepwords=stepwords)
    osr = O.Searcher(O.DB(db, op), op)
    qs = ["".join(rng.choice("ACGT") for _ in range(250)) for _ in range(12)]
    qs += [util.mutate(db[rng.randrange(len(db))][20:270], r, rng) for r in (0.0, 0.02, 0.05, 0.1, 0.2, 0.3) for _ in range(3)]
    qs += ["ACGT" * 60, "A" * 100 + "".join(rng.choice("ACGT") for _ in range(150))]
    n_many = 0
    for K in (33, 200):
        ct, cu, nc, _ = s.rank(qs, K)
        for j, q in enumerate(qs):
            oct, ocu = osr.rank_big(q)
            assert nc[j] == len(oct), (K, j, nc[j], len(oct))
            k = min(K, len(oct))
            assert np.array_equal(ct[j, :k], oct[:k]), (K, j, ct[j, :k][:12], oct[:12])
            assert np.array_equal(cu[j, :k], ocu[:k]), (K, j)
            n_many += len(oct) > 1024
    assert n_many >= 20
