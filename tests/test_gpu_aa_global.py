"""GPU parity of amino acid -usearch_global (BASELINE config 1 and variants) through the C ABI:
k_align<AA> (BLOSUM62, gap open -17, 3-letter HSP words over 20 letters) against the reference
binary's golden files (tools/make_golden_aa_global.py) and against the oracle on fresh inputs."""
import ctypes as C
import os
import random
import subprocess
import sys

import pytest

from tests import util

pytestmark = pytest.mark.gpu


def aa_params(**kw):
    from usearch12_b200 import capi
    p = capi.default_params(**kw)
    capi.lib().usb_set_amino(C.byref(p))
    return p


@pytest.mark.parametrize("variant", list(util.AA_GLOBAL_VARIANTS))
def test_aa_global_matches_reference_golden(variant):
    from usearch12_b200 import capi
    kw = dict(util.AA_GLOBAL_VARIANTS[variant])
    dl, d, ql, q = util.aa_global_inputs(kw.pop("inputs"))
    p = aa_params(**kw)
    ix = capi.Index(d, p, device=0)
    s = capi.Searcher(ix, p)
    res = s.search(q)
    got = util.product_lines(res, ql, q, dl, nucleo=False)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        assert util.first_diff(lines, util.golden_lines(variant, kind)) is None, (variant, kind)
    assert s.launch_count >= 2


def test_aa_global_fresh_inputs_match_oracle():
    """Seeded protein families (indels, substitutions, fragments) at three identity thresholds."""
    from oracle import uso_py as O
    from usearch12_b200 import capi
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import gen_synth_aa
    db, qs = gen_synth_aa.generate(ndb=400, length=350, nq=600, seed=31, nroot=16)
    rng = random.Random(7)
    q = [s for _, s in qs]
    for k in range(0, len(q), 7):   # fragments and extensions: terminal gaps, LA != LB
        a = rng.randrange(0, 120)
        q[k] = q[k][a:a + rng.randrange(60, 250)]
    ql = ["q%d" % i for i in range(len(q))]
    dl = ["p%d" % i for i in range(len(db))]
    for kw in (dict(id=0.5), dict(id=0.8, maxaccepts=2, maxrejects=8), dict(id=0.35, maxaccepts=4, maxrejects=32)):
        p = aa_params(**kw)
        res = capi.Searcher(capi.Index(db, p, device=0), p).search(q)
        got = util.product_lines(res, ql, q, dl, nucleo=False)
        op = O.default_params(amino=True, **kw)
        want = util.oracle_lines(O.Searcher(O.DB(db, op, dl), op), ql, q, dl, nucleo=False)
        for g, w, kind in zip(got, want, ("user", "uc", "b6")):
            assert util.first_diff(g, w) is None, (kw, kind)
        assert len(got[0]) > 100


def test_aa_global_cli_config1(tmp_path):
    """Config 1 through the host CLI: output files byte-identical to the reference binary's."""
    import gzip
    from usearch12_b200 import build
    cli = build.build_cli()
    fa = tmp_path / "test.fa"
    fa.write_bytes(gzip.open(os.path.join(util.GOLDEN, "cfg1_test.fa.gz")).read())
    out = {k: str(tmp_path / k) for k in ("user", "uc", "b6")}
    subprocess.run([cli, "-usearch_global", str(fa), "-db", str(fa), "-id", "0.9", "-uc", out["uc"], "-blast6out", out["b6"],
                    "-userout", out["user"], "-userfields", "query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+caln+qstrand",
                    "-quiet"], check=True)
    for kind in ("user", "uc", "b6"):
        assert open(out[kind]).read().splitlines() == util.golden_lines("cfg1_id90", kind), kind
