"""Dereplication on the device (derepfull.cpp:130-236, seqhash.cpp:6-51,
derepresult.cpp:255-284,689-775): usb_derep_full through the C ABI against a plain restatement, and
-fastx_uniques byte-identical to the reference binary's -fastaout (tools/make_golden_uniques.py)."""
import gzip
import os
import random
import subprocess

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

VARIANTS = {
    "uniq2_sizein": util.UNIQUES2_VARIANTS["sizein"],
    "uniq2_topn": util.UNIQUES2_VARIANTS["topn"],
    "uniq_sizeout": ["-sizeout"],
    "uniq_relabel": ["-sizeout", "-relabel", "Uniq", "-minuniquesize", "2"],
    "uniq_plain": [],
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_fastx_uniques_byte_identical_to_reference(name, tmp_path):
    from usearch12_b200 import build
    cli = build.build_cli()
    src = str(tmp_path / "in.fa")
    with gzip.open(os.path.join(util.GOLDEN, "uniq_in.fa.gz"), "rb") as f, open(src, "wb") as g:
        g.write(f.read())
    dst = str(tmp_path / "out.fa")
    r = subprocess.run([cli, "-fastx_uniques", src, "-fastaout", dst, "-quiet"] + VARIANTS[name], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    with gzip.open(os.path.join(util.GOLDEN, name + ".fa.gz"), "rb") as f:
        want = f.read()
    assert open(dst, "rb").read() == want


def _restated(seqs):
    """DerepFull with -threads 1 (derepfull.cpp:130-212): equal ignoring case, first-occurrence order."""
    seen, out = {}, []
    for s in seqs:
        out.append(seen.setdefault(s.upper(), len(seen)))
    return np.array(out, np.uint32), len(seen)


def test_derep_full_matches_restatement():
    from usearch12_b200 import capi
    rng = random.Random(5)
    base = ["".join(rng.choice("ACGT") for _ in range(rng.choice([1, 2, 31, 32, 33, 100, 250, 251, 1000]))) for _ in range(3000)]
    seqs = []
    for _ in range(40000):
        s = rng.choice(base)
        k = rng.random()
        if k < 0.2:
            s = s.lower()
        elif k < 0.3:
            s = "".join(c.lower() if rng.random() < 0.5 else c for c in s)
        elif k < 0.4:
            s = s[:-1] or "A"           # a prefix: same letters, different length
        elif k < 0.5:
            i = rng.randrange(len(s))
            s = s[:i] + rng.choice("ACGTN") + s[i + 1:]
        seqs.append(s)
    seqs += ["", "A", "a", "", "N" * 64, "n" * 64, "N" * 63]
    got, nu = capi.derep_full(seqs)
    want, wn = _restated(seqs)
    assert nu == wn
    assert np.array_equal(got, want)


def test_derep_full_edge_cases():
    from usearch12_b200 import capi
    got, nu = capi.derep_full([])
    assert nu == 0 and len(got) == 0
    got, nu = capi.derep_full(["ACGT"])
    assert nu == 1 and list(got) == [0]
    got, nu = capi.derep_full(["ACGT"] * 1000)
    assert nu == 1 and not got.any()
    seqs = ["%s" % ("ACGT"[i % 4] * (1 + i % 300)) for i in range(5000)]
    got, nu = capi.derep_full(seqs)
    want, wn = _restated(seqs)
    assert nu == wn and np.array_equal(got, want)
