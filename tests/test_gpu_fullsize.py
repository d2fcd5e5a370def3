"""Size-independent properties at BASELINE.json's full workload (1 M x 250 bp reads vs
100 k x 1500 bp DB, -id 0.97): the oracle cannot run this size in test time, so the hot path is
checked through invariants every reference result satisfies, plus an oracle spot check."""
import os
import sys

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

N_READS = int(os.environ.get("USB_FULL_READS", "1000000"))
N_DB = int(os.environ.get("USB_FULL_DB", "100000"))


@pytest.fixture(scope="module")
def full():
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import synth_np
    from usearch12_b200 import capi
    db, db_off = synth_np.gen_db(N_DB, 1500, seed=4)
    reads, r_off, truth = synth_np.gen_reads(db, db_off, N_READS, 250, seed=1000)
    # plant exact DB windows: every 97th read becomes a verbatim 250-letter window of its target
    planted = np.arange(0, N_READS, 97)
    for i in planted:
        L = int(r_off[i + 1] - r_off[i])
        t = int(truth[i]) if truth[i] >= 0 else int(i % N_DB)
        s = int(db_off[t]) + 100
        reads[int(r_off[i]):int(r_off[i]) + L] = db[s:s + L]
    import ctypes as C
    p = capi.default_params()
    ix = capi.Index.__new__(capi.Index)
    ix.params, ix._data, ix._off, ix.n_seq = p, db, db_off, N_DB
    h = C.c_void_p()
    capi.check(capi.lib().usb_index_create(0, C.byref(p), db.ctypes.data_as(C.c_void_p), db_off.ctypes.data_as(C.c_void_p),
                                           N_DB, C.byref(h)))
    ix.handle = h
    s = capi.Searcher(ix, p)
    res = s.search_packed(reads, r_off)
    return dict(capi=capi, db=db, db_off=db_off, reads=reads, r_off=r_off, planted=planted, s=s, res=res, p=p)


def test_every_hit_is_a_valid_accepted_alignment(full):
    res, r_off = full["res"], full["r_off"]
    h = res.hits
    assert len(h) > 0.6 * N_READS
    ql = (r_off[1:] - r_off[:-1]).astype(np.int64)
    assert np.array_equal(h["ql"], ql[h["query"]])
    assert (h["target"] < N_DB).all()
    # accepter.cpp:36-38: accepted <=> not (ids/cols < (double)(float)0.97)
    frac = h["ids"].astype(np.float64) / h["alnlen"].astype(np.float64)
    assert (frac >= np.float64(np.float32(0.97))).all()
    assert (h["ids"] + h["mism"] + h["intgaps"] == h["alnlen"]).all()
    assert (h["opens"] <= h["intgaps"]).all()
    # paths: run lengths give back both sequence lengths
    runs = res.runs
    op = runs & 3
    ln = (runs >> 2).astype(np.int64)
    start = h["run_off"].astype(np.int64)
    cnt = h["run_cnt"].astype(np.int64)
    seg = np.repeat(np.arange(len(h)), cnt)
    idx = np.concatenate([np.arange(s, s + c) for s, c in zip(start[:2000], cnt[:2000])]) if len(h) else np.zeros(0, int)
    seg = seg[:len(idx)]
    m = np.bincount(seg, weights=ln[idx] * (op[idx] == 0), minlength=2000)
    d = np.bincount(seg, weights=ln[idx] * (op[idx] == 1), minlength=2000)
    i = np.bincount(seg, weights=ln[idx] * (op[idx] == 2), minlength=2000)
    n = min(2000, len(h))
    assert np.array_equal((m + d)[:n], h["ql"][:n]) and np.array_equal((m + i)[:n], h["tl"][:n])
    # one accept per query at most (-maxaccepts 1), queries ascending
    assert (np.diff(h["query"].astype(np.int64)) > 0).all()


def test_planted_exact_windows_hit_at_100_percent(full):
    res, planted = full["res"], full["planted"]
    qoff = res.qoff.astype(np.int64)
    has = qoff[planted + 1] > qoff[planted]
    assert has.mean() > 0.999
    hh = res.hits[qoff[planted[has]]]
    assert (hh["ids"] == hh["alnlen"]).all() and (hh["alnlen"] == hh["ql"]).all() and (hh["mism"] == 0).all()


def test_batch_split_invariance_and_idempotence(full):
    capi, s, res = full["capi"], full["s"], full["res"]
    reads, r_off = full["reads"], full["r_off"]
    cut = N_READS // 3
    a = s.search_packed(reads, r_off[:cut + 1])
    b = s.search_packed(reads, r_off[cut:])
    key = ["target", "ids", "mism", "intgaps", "alnlen", "first_mq", "last_mt", "rank"]
    merged = np.concatenate([a.hits[key], b.hits[key]])
    assert np.array_equal(merged, res.hits[key])
    q = np.concatenate([a.hits["query"], b.hits["query"] + cut])
    assert np.array_equal(q, res.hits["query"])
    again = s.search_packed(reads, r_off)
    assert np.array_equal(again.hits[key], res.hits[key])


def test_oracle_spot_check_on_the_full_database(full):
    """500 of the reads, searched by the oracle against the same 100 k-target DB."""
    from oracle import uso_py as O
    db, db_off, reads, r_off, res = full["db"], full["db_off"], full["reads"], full["r_off"], full["res"]
    raw = db.tobytes()
    seqs = [raw[int(db_off[i]):int(db_off[i + 1])] for i in range(N_DB)]
    op = O.default_params()
    osr = O.Searcher(O.DB(seqs, op), op)
    rraw = reads.tobytes()
    qoff = res.qoff.astype(np.int64)
    rng = np.random.default_rng(1)
    for qi in rng.choice(N_READS, size=500, replace=False):
        q = rraw[int(r_off[qi]):int(r_off[qi + 1])]
        want = osr.search(q, int(qi))
        b, e = int(qoff[qi]), int(qoff[qi + 1])
        assert e - b == len(want), qi
        for k, w in enumerate(want):
            h = res.hits[b + k]
            assert int(h["target"]) == w["target"] and int(h["ids"]) == w["ids"] and res.path(h) == w["path"], qi


def test_first_20000_reads_identical_to_the_reference_binary(full, tmp_path):
    """The unmodified reference binary (oracle/_ref/usearch12, built by oracle/Makefile.ref; it travels
    with the repository) searches the first 20 000 reads against the same 100 k-target database; the
    command line of this repository must write the same .uc and .b6 bytes (search.cpp:63-86 ->
    outputuc.cpp:45-93, blast6out.cpp:27-80)."""
    import subprocess
    import synth_np
    from usearch12_b200 import build
    ref = os.path.join(util.ROOT, "oracle", "_ref", "usearch12")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/usearch12 not built (python __graft_entry__.py builds it where /root/reference exists)")
    n = min(20000, N_READS)
    dbfa, qfa = str(tmp_path / "db.fa"), str(tmp_path / "q.fa")
    synth_np.write_fasta(dbfa, full["db"], full["db_off"], "db")
    synth_np.write_fasta(qfa, full["reads"], full["r_off"], "q", 0, n)
    out = {}
    for name, exe in (("ref", ref), ("usb", build.build_cli())):
        uc, b6 = str(tmp_path / (name + ".uc")), str(tmp_path / (name + ".b6"))
        cmd = [exe, "-usearch_global", qfa, "-db", dbfa, "-id", "0.97", "-strand", "plus", "-uc", uc, "-blast6out", b6, "-quiet"]
        if name == "ref":
            cmd += ["-threads", "1"]  # one thread: the binary writes hits in the order its threads finish
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
        out[name] = (open(uc, "rb").read(), open(b6, "rb").read())
    assert len(out["ref"][0].splitlines()) == n
    assert out["usb"][0] == out["ref"][0], "uc"
    assert out["usb"][1] == out["ref"][1], "b6"
    assert len(out["ref"][1]) > 0
