"""The whole command on the device with every output file: usearch12_b200_cli (FASTA in, kernels, HitMgr
mirror, OutputSink / DBHitSink) against the files the unmodified reference binary wrote for the same command
line (tests/golden/fmt_*, tools/make_golden_formats.py).  The formats alone are pinned without a GPU by
tests/test_formats_cpu.py; here the hits come from the CUDA path and the database letters from the index."""
import os
import subprocess
import sys

import pytest

from tests import util
from tests.test_formats_cpu import OUT_FLAGS, check_outputs, out_paths

sys.path.insert(0, os.path.join(util.ROOT, "tools"))
import make_golden_formats as M  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(M.VARIANTS))
def test_cli_writes_the_reference_files(name, tmp_path):
    from usearch12_b200 import build
    cli = build.build_cli()
    tmp = str(tmp_path)
    q, d = M.write_inputs(name, tmp)
    cmd_name, opts, fields = M.VARIANTS[name][0], M.VARIANTS[name][4], M.VARIANTS[name][5]
    paths = out_paths(name, tmp)
    cmd = [cli, "-" + cmd_name, q, "-db", d, "-quiet", "-userfields", fields] + opts
    for k, path in paths.items():
        cmd += [OUT_FLAGS[k], path]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    check_outputs(name, paths)


def test_cli_fastq_queries(tmp_path):
    """FASTQ reads as queries: same hits as the FASTA form, -matchedfq / -notmatchedfq as the reference writes them."""
    from usearch12_b200 import build
    from tests.test_formats_cpu import golden_bytes
    cli = build.build_cli()
    tmp = str(tmp_path)
    _, d = M.write_inputs("fmt_nt", tmp)
    q = os.path.join(tmp, "q.fq")
    M.write_fastq(q)
    outs = {k: os.path.join(tmp, "o." + k) for k in ("matchedfq", "notmatchedfq", "matched", "uc")}
    cmd = [cli, "-usearch_global", q, "-db", d, "-quiet"] + M.VARIANTS["fmt_nt"][4]
    for k, path in outs.items():
        cmd += ["-" + k, path]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for k, path in outs.items():
        assert open(path, "rb").read() == golden_bytes("fmt_fq" if k.endswith("fq") else "fmt_nt", k), k


@pytest.mark.parametrize("name", ["sc_a", "sc_b", "sc_c", "sc_d", "sc_e"])
def test_score_options_match_reference(name, tmp_path):
    """-match / -mismatch / -minhsp / -xdrop_nw: the C ABI with the options in usb_params and the CLI with the
    reference's option names, against the reference binary's files (tools/make_golden_scores.py)."""
    import make_golden_scores as S
    from usearch12_b200 import build, capi
    recs = M.query_subset("fmt_nt")
    qlab, qs = [r[0][1:] for r in recs], [r[1] for r in recs]
    dlab, db = util.read_fasta(os.path.join(util.GOLDEN, "db.fa.gz"))
    g = util.Golden()
    p = capi.default_params(**dict(S.PARAMS, **S.VARIANTS[name][1]))
    s = capi.Searcher(capi.Index(db, p), p)
    got = util.product_lines(s.search(qs), qlab, qs, dlab)
    for lines, kind in zip(got, ("user", "uc", "b6")):
        d = util.first_diff(lines, g.lines(name, kind))
        assert d is None, "C ABI %s %s\n%s" % (name, kind, d)
    tmp = str(tmp_path)
    q, d = M.write_inputs("fmt_nt", tmp)
    outs = {k: os.path.join(tmp, "o." + k) for k in ("user", "uc", "b6")}
    cmd = [build.build_cli(), "-usearch_global", q, "-db", d, "-quiet"] + S.BASE + S.VARIANTS[name][0] + [
        "-userout", outs["user"], "-userfields", S.USERFIELDS, "-uc", outs["uc"], "-blast6out", outs["b6"]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind, path in outs.items():
        dd = util.first_diff(open(path).read().splitlines(), g.lines(name, kind))
        assert dd is None, "CLI %s %s\n%s" % (name, kind, dd)


@pytest.mark.parametrize("name", ["exh_00", "exh_0r", "exh_a0"])
def test_exhaustive_search_matches_reference(name, tmp_path):
    """-maxaccepts 0 / -maxrejects 0 against 3 000 targets (tools/make_golden_exhaustive.py): the candidate loop
    runs over whole U-sorted lists made by k_usort_full (up to 2 331 candidates here, more than the 1 024 k_rank
    materialises); hits identical to the reference binary's, through the C ABI and through the CLI in small
    batches."""
    import make_golden_exhaustive as X
    from usearch12_b200 import build, capi
    db, dlab, qs, qlab = X.inputs()
    g = util.Golden()
    p = capi.default_params(**X.VARIANTS[name][1])
    s = capi.Searcher(capi.Index(db, p), p)
    res = s.search(qs)
    got = util.product_lines(res, qlab, qs, dlab)
    for lines, kind in zip(got[:2], ("user", "uc")):
        d = util.first_diff(lines, g.lines(name, kind))
        assert d is None, "C ABI %s %s\n%s" % (name, kind, d)
    assert int(res.qstat["n_cand"].max()) > 1024
    if name != "exh_00":
        return
    tmp = str(tmp_path)
    q, d = os.path.join(tmp, "q.fa"), os.path.join(tmp, "db.fa")
    open(q, "w").write("".join(">%s\n%s\n" % x for x in zip(qlab, qs)))
    open(d, "w").write("".join(">%s\n%s\n" % x for x in zip(dlab, db)))
    outs = {k: os.path.join(tmp, "o." + k) for k in ("user", "uc")}
    cmd = [build.build_cli(), "-usearch_global", q, "-db", d, "-quiet", "-batch", "100"] + X.VARIANTS[name][0] + [
        "-userout", outs["user"], "-userfields", X.USERFIELDS, "-uc", outs["uc"]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind, path in outs.items():
        dd = util.first_diff(open(path).read().splitlines(), g.lines(name, kind))
        assert dd is None, "CLI %s %s\n%s" % (name, kind, dd)


def test_exhaustive_local_search_matches_reference():
    """-usearch_local -maxaccepts 0 -maxrejects 0: k_local walks the whole candidate lists of k_usort_full."""
    import make_golden_exhaustive as X
    from usearch12_b200 import capi
    db, dlab, qs, qlab = X.inputs()
    name, _, kw, ev = X.LOCAL
    p = util.product_local_params(True, ev, **kw)
    s = capi.Searcher(capi.Index(db, p), p)
    res = s.search(qs)
    got = util.product_lines_local(res, s, qlab, qs, dlab, True)
    g = util.Golden()
    for lines, kind in zip(got[:2], ("user", "uc")):
        d = util.first_diff(lines, g.lines(name, kind))
        assert d is None, "%s %s\n%s" % (name, kind, d)
    assert int(res.qstat["n_cand"].max()) > 1024


def test_exhaustive_amino_search_matches_reference():
    """Amino acid -usearch_global -maxaccepts 0 -maxrejects 0 on 1 500 proteins: k_align takes its candidates
    from k_usort_full's lists (stride of N candidates per query)."""
    import ctypes as C
    import make_golden_exhaustive as X
    from usearch12_b200 import capi
    db, dlab, qs, qlab = X.inputs_aa()
    name, _, kw = X.AMINO
    p = capi.default_params(**kw)
    capi.lib().usb_set_amino(C.byref(p))
    s = capi.Searcher(capi.Index(db, p), p)
    got = util.product_lines(s.search(qs), qlab, qs, dlab, nucleo=False)
    g = util.Golden()
    for lines, kind in zip(got[:2], ("user", "uc")):
        d = util.first_diff(lines, g.lines(name, kind))
        assert d is None, "%s %s\n%s" % (name, kind, d)


@pytest.mark.parametrize("name", ["exh_qt5", "exh_qt3"])
def test_skipped_pairs_walk_whole_candidate_lists(name, tmp_path):
    """RejectPair rules skip pairs without a Terminator call (searcher.cpp:63-67).  With -minqt 0.5 every pair of
    these inputs is skipped and two queries walk lists of 1 363 and 2 331 candidates: the batch runs out of the
    1 024 candidates k_rank materialises and is repeated with the whole lists (k_usort_full).  Files identical to
    the reference binary's (tools/make_golden_exhaustive.py)."""
    import make_golden_exhaustive as X
    from usearch12_b200 import build
    db, dlab, qs, qlab = X.inputs()
    g = util.Golden()
    tmp = str(tmp_path)
    q, d = os.path.join(tmp, "q.fa"), os.path.join(tmp, "db.fa")
    open(q, "w").write("".join(">%s\n%s\n" % x for x in zip(qlab, qs)))
    open(d, "w").write("".join(">%s\n%s\n" % x for x in zip(dlab, db)))
    outs = {k: os.path.join(tmp, "o." + k) for k in ("user", "uc")}
    cmd = [build.build_cli(), "-usearch_global", q, "-db", d, "-quiet"] + X.SKIPS[name] + [
        "-userout", outs["user"], "-userfields", X.USERFIELDS, "-uc", outs["uc"]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind, path in outs.items():
        dd = util.first_diff(open(path).read().splitlines(), g.lines(name, kind))
        assert dd is None, "%s %s\n%s" % (name, kind, dd)


def test_cli_otutab_with_biom(tmp_path):
    """-otutab with -biomout through the CLI on the device (the table, the map and the BIOM file of the reference)."""
    from usearch12_b200 import build
    from tests.test_formats_cpu import _otutab_inputs, check_otutab_outputs
    tmp = str(tmp_path)
    _otutab_inputs(tmp)
    r = subprocess.run([build.build_cli(), "-otutab", "otutab_reads.fa", "-otus", "otutab_otus.fa", "-otutabout", "tab.txt",
                        "-mapout", "map.txt", "-biomout", "o.biom", "-dbmatched", "dbm.fa", "-dbnotmatched", "dbnm.fa",
                        "-notmatched", "nm.fa", "-sizeout", "-quiet"], cwd=tmp, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    check_otutab_outputs(tmp)


def test_cluster_fast_writes_the_output_sink_files(tmp_path):
    """-cluster_fast with the per-hit files of its OutputSink (-userout, -blast6out, -alnout, -fastapairs, -matched,
    -notmatched) next to the .uc: the reference binary's files (tools/make_golden_cluster_outputs.py)."""
    import gzip
    import json
    import make_golden_cluster_outputs as C
    from usearch12_b200 import build
    tmp = str(tmp_path)
    reads = os.path.join(tmp, "r.fa")
    with gzip.open(os.path.join(util.GOLDEN, "cluster_reads.fa.gz"), "rb") as f, open(reads, "wb") as g:
        g.write(f.read())
    outs = {k: os.path.join(tmp, "o." + k) for k in C.FLAGS}
    cmd = [build.build_cli(), "-cluster_fast", reads, "-quiet", "-userfields", C.USERFIELDS] + C.OPTS
    for k, flag in C.FLAGS.items():
        cmd += [flag, outs[k]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    g = util.Golden()
    for kind in ("user", "b6"):
        d = util.first_diff(open(outs[kind]).read().splitlines(), g.lines("cluster_out", kind))
        assert d is None, "%s\n%s" % (kind, d)
    d = util.first_diff(open(outs["uc"]).read().splitlines(), g.lines("cluster_length", "uc"))
    assert d is None, d
    sums = json.load(open(os.path.join(util.GOLDEN, "cluster_out_sha256.json")))
    for kind, want in sums.items():
        assert C.digest_of(kind, open(outs[kind], "rb").read()) == want, kind
