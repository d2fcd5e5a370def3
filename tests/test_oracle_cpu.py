"""CPU-only: the oracle restatement against the golden files written by the UNMODIFIED reference
binary (tools/make_golden.py), plus host-logic checks that need no GPU."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import uso_py as O
from tests import util


@pytest.mark.parametrize("variant", ["plus97", "both80_ma0"])
def test_oracle_matches_reference_golden(golden, variant):
    kw = util.VARIANTS[variant]
    p = O.default_params(**kw)
    db = O.DB(golden.db, p, golden.db_labels)
    s = O.Searcher(db, p)
    # a slice that covers every edge-case block of tools/make_golden.py (the tail) plus regular reads
    idx = list(range(0, 300)) + list(range(2400, len(golden.q)))
    want = {k: golden.lines(variant, k) for k in ("user", "uc", "b6")}
    labels = [golden.q_labels[i] for i in idx]
    seqs = [golden.q[i] for i in idx]
    user, uc, b6 = util.oracle_lines(s, labels, seqs, golden.db_labels)
    keep = set(labels)
    for got, kind, col in ((user, "user", 0), (uc, "uc", 8), (b6, "b6", 0)):
        ref = [l for l in want[kind] if l.split("\t")[col] in keep]
        assert util.first_diff(got, ref) is None, kind


@pytest.mark.parametrize("variant", list(util.LOCAL_VARIANTS))
def test_oracle_local_matches_reference_golden(variant):
    """usearch_local restatement vs the reference binary's outputs (tools/make_golden_local.py)."""
    kw = dict(util.LOCAL_VARIANTS[variant])
    nucleo = kw.pop("nucleo")
    g = util.GoldenLocal("nt" if nucleo else "aa")
    p = util.oracle_local_params(nucleo, **kw)
    s = O.Searcher(O.DB(g.db, p, g.db_labels), p)
    got = util.oracle_lines_local(s, g.q_labels, g.q, g.db_labels, nucleo)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        assert util.first_diff(lines, g.lines(variant, kind)) is None, kind


@pytest.mark.parametrize("variant", list(util.LOCAL_LONG_VARIANTS))
def test_oracle_local_split_extensions_match_reference_golden(variant):
    """Sequences above g_MaxL = 4096 letters: XDropFwdSplit / XDropBwdSplit (xdropfwdsplit.cpp:24-91,
    xdropbwdsplit.cpp:15-79) vs the reference binary's outputs (tools/make_golden_local_long.py)."""
    kw = dict(util.LOCAL_LONG_VARIANTS[variant])
    nucleo = kw.pop("nucleo")
    g = util.GoldenLocal("nt" if nucleo else "aa", prefix="loclong")
    p = util.oracle_local_params(nucleo, **kw)
    s = O.Searcher(O.DB(g.db, p, g.db_labels), p)
    got = util.oracle_lines_local(s, g.q_labels, g.q, g.db_labels, nucleo)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        assert util.first_diff(lines, g.lines(variant, kind)) is None, kind
    assert max(int(l.split("\t")[3]) for l in got[0]) > 4096


@pytest.mark.parametrize("variant", list(util.AA_GLOBAL_VARIANTS))
def test_oracle_aa_global_matches_reference_golden(variant):
    """Amino acid usearch_global (BASELINE config 1 = cfg1_id90) vs the reference binary's outputs
    (tools/make_golden_aa_global.py): BLOSUM62, gap open -17, HSP words of 3 letters."""
    kw = dict(util.AA_GLOBAL_VARIANTS[variant])
    dl, d, ql, q = util.aa_global_inputs(kw.pop("inputs"))
    p = O.default_params(amino=True, **kw)
    s = O.Searcher(O.DB(d, p, dl), p)
    got = util.oracle_lines(s, ql, q, dl, nucleo=False)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        assert util.first_diff(lines, util.golden_lines(variant, kind)) is None, kind


def test_known_answer_xdrop_fwd():
    """The reference's own known answer (xdropalignmem.cpp:336-364 cmd_test): XDropFwdFastMem of
    SEQVENCE / SEQVECE with BLOSUM62 scores 27.0, Leni 8, Lenj 7, alignment SEQVENCE / SEQVE-CE."""
    p = O.default_params(amino=True, local=1)
    assert O.xdrop_fwd(p, "SEQVENCE", "SEQVECE", 32.0) == (27.0, 8, 7, "MMMMMDMM")


def test_known_answer_compress_path():
    assert O.compress_path("I" * 610 + "M" * 250 + "I" * 638) == "610I250M638I"
    assert O.compress_path("I" * 847 + "M" * 12 + "D" + "M" * 238 + "I" * 408) == "847I12MD238M408I"


def test_capi_library_exports_every_symbol():
    from usearch12_b200 import capi
    L = capi.lib()
    for name in capi.SYMBOLS:
        assert hasattr(L, name), name
    # header and binding agree on the struct sizes
    p = capi.default_params()
    assert p.struct_size == C.sizeof(capi.Params)
    assert capi.HIT_DTYPE.itemsize == 80 and capi.QSTAT_DTYPE.itemsize == 28


def test_header_declares_bound_symbols():
    import os
    import re
    hdr = open(os.path.join(util.ROOT, "include", "usb200.h")).read()
    from usearch12_b200 import capi
    declared = set(re.findall(r"\b(usb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_no_gpu_fails_loudly():
    from usearch12_b200 import capi
    if capi.lib().usb_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(capi.UsbError):
        capi.Index(["ACGTACGTACGTACGT"])


def test_product_never_imports_oracle():
    import os
    for root, _, files in os.walk(os.path.join(util.ROOT, "usearch12_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, fn), errors="ignore").read()
                assert "uso_" not in txt and "liboracle" not in txt, fn


def test_oracle_cluster_fast_matches_reference_golden(tmp_path):
    """oracle/uso_cli cluster_fast vs the reference binary's .uc / centroids (tests/golden/cluster_*)."""
    import gzip
    import os
    import subprocess
    O.lib()
    cli = os.path.join(util.ROOT, "oracle", "_build", "uso_cli")
    reads = os.path.join(str(tmp_path), "g.fa")
    with gzip.open(os.path.join(util.GOLDEN, "cluster_reads.fa.gz"), "rb") as fi, open(reads, "wb") as fo:
        fo.write(fi.read())
    for sort in ("none", "size"):
        uc, cen = os.path.join(str(tmp_path), "o.uc"), os.path.join(str(tmp_path), "o.fa")
        subprocess.run([cli, "cluster_fast", reads, "0.97", uc, cen, sort], check=True)
        for got, name in ((uc, "cluster_%s.uc.gz" % sort), (cen, "cluster_%s.centroids.fa.gz" % sort)):
            with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
                assert util.first_diff(open(got).read().splitlines(), f.read().splitlines()) is None, name


@pytest.mark.parametrize("name", ["sc_a", "sc_b", "sc_c", "sc_d", "sc_e"])
def test_oracle_score_options_match_reference_golden(name):
    """-match / -mismatch / -minhsp / -xdrop_nw (alnparams.cpp:330-334, alnheuristics.cpp:26-44): the oracle
    with the same options against the reference binary's files (tools/make_golden_scores.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import make_golden_formats as M
    import make_golden_scores as S
    from oracle import uso_py as O
    recs = M.query_subset("fmt_nt")
    qlab, qs = [r[0][1:] for r in recs], [r[1] for r in recs]
    dlab, db = util.read_fasta(os.path.join(util.GOLDEN, "db.fa.gz"))
    op = O.default_params(**dict(S.PARAMS, **S.VARIANTS[name][1]))
    want = util.oracle_lines(O.Searcher(O.DB(db, op, dlab), op), qlab, qs, dlab)
    g = util.Golden()
    for got, kind in zip(want, ("user", "uc", "b6")):
        d = util.first_diff(got, g.lines(name, kind))
        assert d is None, "%s %s\n%s" % (name, kind, d)


@pytest.mark.parametrize("name", ["exh_00", "exh_0r", "exh_a0"])
def test_oracle_exhaustive_search_matches_reference_golden(name):
    """-maxaccepts 0 / -maxrejects 0 (terminator.cpp:23-31) on 3 000 targets: the oracle's candidate loop runs
    through whole U-sorted lists (up to ~2 500 candidates for the random reads) like the reference binary
    (tools/make_golden_exhaustive.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import make_golden_exhaustive as X
    from oracle import uso_py as O
    db, dlab, qs, qlab = X.inputs()
    op = O.default_params(**X.VARIANTS[name][1])
    srch = O.Searcher(O.DB(db, op, dlab), op)
    want = util.oracle_lines(srch, qlab, qs, dlab)
    g = util.Golden()
    for got, kind in zip(want[:2], ("user", "uc")):
        d = util.first_diff(got, g.lines(name, kind))
        assert d is None, "%s %s\n%s" % (name, kind, d)


def test_oracle_exhaustive_local_search_matches_reference_golden():
    """-usearch_local -maxaccepts 0 -maxrejects 0 on the same 3 000 targets (tests/golden/exh_loc.*)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import make_golden_exhaustive as X
    from oracle import uso_py as O
    db, dlab, qs, qlab = X.inputs()
    name, _, kw, ev = X.LOCAL
    op = util.oracle_local_params(True, evalue=ev, **kw)
    want = util.oracle_lines_local(O.Searcher(O.DB(db, op, dlab), op), qlab, qs, dlab, True)
    g = util.Golden()
    for got, kind in zip(want[:2], ("user", "uc")):
        d = util.first_diff(got, g.lines(name, kind))
        assert d is None, "%s %s\n%s" % (name, kind, d)


def test_oracle_exhaustive_amino_search_matches_reference_golden():
    """Amino acid -usearch_global -maxaccepts 0 -maxrejects 0 on 1 500 proteins (tests/golden/exh_aa.*)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import make_golden_exhaustive as X
    from oracle import uso_py as O
    db, dlab, qs, qlab = X.inputs_aa()
    name, _, kw = X.AMINO
    op = O.default_params(amino=True, **kw)
    want = util.oracle_lines(O.Searcher(O.DB(db, op, dlab), op), qlab, qs, dlab, nucleo=False)
    g = util.Golden()
    for got, kind in zip(want[:2], ("user", "uc")):
        d = util.first_diff(got, g.lines(name, kind))
        assert d is None, "%s %s\n%s" % (name, kind, d)
