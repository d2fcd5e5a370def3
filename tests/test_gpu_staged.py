"""The staged candidate loop (k_gate -> k_dp -> k_commit, usb_stage.cuh) against the one-kernel
loop (k_align) and the oracle: identical hits, paths and per-query counters for every
Terminator setting, both strands, wildcards, -fulldp and -band 0."""
import os
import random

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _search(golden, qs, **kw):
    from usearch12_b200 import capi
    p = capi.default_params(**kw)
    ix = capi.Index(golden.db, p, device=0)
    s = capi.Searcher(ix, p)
    res = s.search(qs)
    return res, s


@pytest.mark.parametrize("kw", [
    dict(id=0.97), dict(id=0.9, strand_both=1, maxaccepts=4, maxrejects=64), dict(id=0.8, maxaccepts=3, maxrejects=16),
    dict(id=0.5, maxaccepts=0, maxrejects=0), dict(id=0.9, maxaccepts=2, maxrejects=0), dict(id=0.97, fulldp=1),
    dict(id=0.9, band=0)])
def test_staged_equals_one_kernel(golden, kw):
    qs = golden.q[:700] + golden.q[2400:]
    os.environ.pop("USB_ONE_KERNEL_ALIGN", None)
    a, sa = _search(golden, qs, **kw)
    assert sa.kernel_ms()["gate"] > 0.0          # the staged pipeline ran
    os.environ["USB_ONE_KERNEL_ALIGN"] = "1"
    try:
        b, sb = _search(golden, qs, **kw)
        assert sb.kernel_ms()["gate"] == 0.0
    finally:
        os.environ.pop("USB_ONE_KERNEL_ALIGN", None)
    assert np.array_equal(a.qoff, b.qoff)
    assert len(a.hits) == len(b.hits) and len(a.hits) > 100
    for f in ("query", "target", "strand", "rank", "ids", "mism", "intgaps", "opens", "alnlen", "first_mq", "first_mt",
              "last_mq", "last_mt", "first_mcol", "ql", "tl"):
        assert np.array_equal(a.hits[f], b.hits[f]), f
    for i in range(len(a.hits)):
        assert a.path(a.hits[i]) == b.path(b.hits[i]), i
    for f in ("n_cand", "n_tried", "n_hspfail", "n_dp", "dp_cells", "n_accept", "seq_bytes"):
        assert np.array_equal(a.qstat[f], b.qstat[f]), f


def test_staged_many_stages_matches_oracle():
    """-maxaccepts 0 -maxrejects 0 on a DB of 300 targets: every candidate is examined, which takes
    several stages of 64 candidates; hits in HitMgr order equal to the oracle's."""
    import sys
    from oracle import uso_py as O
    from usearch12_b200 import capi
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    from gen_synth import generate
    db, reads = generate(ndb=300, dblen=700, nq=150, qlen=200, seed=77, nroot=2)
    qs = [r[1] for r in reads]
    ql = [r[0] for r in reads]
    dl = ["db%d" % i for i in range(len(db))]
    kw = dict(id=0.8, maxaccepts=0, maxrejects=0)
    p = capi.default_params(**kw)
    s = capi.Searcher(capi.Index(db, p, device=0), p)
    res = s.search(qs)
    assert s.kernel_ms()["gate"] > 0.0
    op = O.default_params(**kw)
    want = util.oracle_lines(O.Searcher(O.DB(db, op, dl), op), ql, qs, dl)
    got = util.product_lines(res, ql, qs, dl)
    for g, w, kind in zip(got, want, ("user", "uc", "b6")):
        assert util.first_diff(g, w) is None, kind
    assert len(got[0]) > 2000
