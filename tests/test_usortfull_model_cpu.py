"""The algorithm of k_usort_full (usearch12_b200/csrc/usb_usortfull.cuh) restated step for step in Python -- 32
targets per step, the fast path for steps without a new running maximum, NextValue taken from the running maximum,
histogram, descending offsets, stable placement -- against the oracle's SetTopBump + CountSortOrderDesc
(udbusortedsearcher.cpp:205-267, countsort.cpp:6-108) on real U vectors, for -bump 0 / 50 / 80 / 100 and databases
with many ties.  This checks the claims the kernel's shortcuts rest on; the kernel itself is compared with the
reference binary's files on the device (tests/test_gpu_zz_formats.py)."""
import os
import random
import sys

import numpy as np
import pytest

from tests import util

sys.path.insert(0, os.path.join(util.ROOT, "tools"))


def usort_full_model(U, bump_pct):
    N = len(U)
    bump_d = bump_pct / 100.0
    min_u, max_count, next_value = 1, 0, 0
    surv = []
    for base in range(0, N, 32):                       # pass A
        u = [int(x) for x in U[base:base + 32]]
        if not any(x > max_count for x in u):          # no new maximum in this step: MinU cannot change
            keep = [x >= min_u for x in u]
        else:                                          # replayed lane by lane
            keep = []
            for n in u:
                k = False
                if n >= min_u:
                    if n > max_count:
                        if bump_d != 0.0:
                            new = int(n * bump_d)
                            if new > min_u and new < max_count:
                                min_u = new
                        next_value = max_count
                        max_count = n
                    k = True
                keep.append(k)
        surv += [base + i for i, k in enumerate(keep) if k]
    min_value = next_value // 2
    hist = [0] * (max_count + 1)                       # pass B
    for t in surv:
        if U[t] >= min_value:
            hist[int(U[t])] += 1
    total = 0
    hi = max_count + 1
    while hi > min_value:                              # descending offsets, 32 values per step
        n = min(32, hi - min_value)
        vals = [hi - 1 - lane for lane in range(n)]
        inc = 0
        for v in vals:
            c = hist[v]
            hist[v] = total + inc
            inc += c
        total += inc
        hi -= n
    out = [None] * total                               # pass C
    for i in range(0, len(surv), 32):
        chunk = [t for t in surv[i:i + 32]]
        groups = {}
        for t in chunk:
            if U[t] >= min_value:
                groups.setdefault(int(U[t]), []).append(t)
        for v, ts in groups.items():
            slot = hist[v]
            hist[v] = slot + len(ts)
            for r, t in enumerate(ts):
                out[slot + r] = t
    return out


@pytest.mark.parametrize("bump", [0, 50, 80, 100])
@pytest.mark.parametrize("shape", ["families", "ties", "tiny"])
def test_model_of_k_usort_full_equals_the_oracle(bump, shape):
    from gen_synth import generate
    from oracle import uso_py as O
    rng = random.Random(bump * 7 + len(shape))
    if shape == "families":
        db, reads = generate(ndb=1400, dblen=500, nq=40, qlen=200, seed=3 + bump, nroot=14)
    elif shape == "ties":  # two families of near copies: long runs of equal counters
        db, reads = generate(ndb=700, dblen=400, nq=40, qlen=150, seed=5 + bump, nroot=2)
        db = [s if i % 3 else db[i % 2] for i, s in enumerate(db)]
    else:  # fewer targets than one step, and one more than a step
        db, reads = generate(ndb=33, dblen=300, nq=30, qlen=120, seed=9 + bump, nroot=3)
    qs = [r[1] for r in reads] + ["".join(rng.choice("ACGT") for _ in range(200)) for _ in range(6)]
    op = O.default_params(bump=bump, maxaccepts=0, maxrejects=0)
    s = O.Searcher(O.DB(db, op), op)
    longest = 0
    for q in qs:
        U, ct, _ = s.rank(q)
        got = usort_full_model(U, bump)
        assert got == [int(t) for t in ct]
        longest = max(longest, len(got))
    assert longest > (100 if shape != "tiny" else 10)
