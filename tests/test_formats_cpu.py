"""Output formats of the host mirror (OutputSink / DBHitSink in usearch12_b200/csrc/host) against files
written by the unmodified reference binary, WITHOUT a GPU: tools/format_replay.cpp feeds the reference's
own hit table (tests/golden/fmt_*.hits.gz = its -userout with the fields of a usb_hit) to the sinks, and
every file they write must be the reference's byte for byte: -alnout (alnout.cpp:45-171 and the per-query
report outputsink.cpp:237-356), -fastapairs / -qsegout / -tsegout (outputsink.cpp:17-44,197-235), -matched /
-notmatched (:383-400), -dbmatched / -dbnotmatched (dbhitsink.cpp:108-159), -uc, -blast6out, and -userout
with 66 userfields (userout.cpp:126-352).  Fixtures: tools/make_golden_formats.py."""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import pytest

from tests import util

sys.path.insert(0, os.path.join(util.ROOT, "tools"))
import make_golden_formats as M  # noqa: E402

REPLAY_OPTS = {
    "fmt_nt": [],
    "fmt_nh": ["-output_no_hits"],
    "fmt_ms": ["-minsize", "5", "-uc_hitsonly"],
    "fmt_rl": ["-rowlen", "50", "-flank", "3"],
    "fmt_sz": ["-sizein", "-sizeout"],
    "fmt_aag": ["-amino", "1"],
    "fmt_aal": ["-amino", "1", "-local", "1", "-evalue", "10"],
    "fmt_ntl": ["-local", "1", "-evalue", "1e-3"],
}
OUT_FLAGS = dict(M.FLAGS, user="-userout")


def out_paths(name, tmp):
    """kind -> output file for the kinds a variant stores (plus the two database files when it stores them)."""
    kinds = [k for k in M.kinds_of(name) if k != "hits"]
    if len(M.VARIANTS[name]) == 6:
        kinds += ["dbm", "dbnm", "dbcut"]
    return {k: os.path.join(tmp, "o." + k) for k in kinds}


def golden_bytes(name, kind):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.gz" % (name, kind)), "rb") as f:
        return f.read()


def check_outputs(name, paths):
    """paths: kind -> file written by the product; compares with the reference's files."""
    sums = json.load(open(os.path.join(util.GOLDEN, "fmt_db_sha256.json")))
    for kind, path in paths.items():
        got = open(path, "rb").read()
        if kind in ("dbm", "dbnm", "dbcut"):
            want = sums["%s.%s" % (name, kind)]
            assert (len(got), got.count(b">")) == (want["bytes"], want["seqs"]), (name, kind)
            assert hashlib.sha256(got).hexdigest() == want["sha256"], (name, kind)
            continue
        if kind == "aln":  # line 1 = command line, line 2 = program / host line
            head = got.split(b"\n")[:2]
            assert len(head) == 2 and head[0] and head[1], "alnout header lines"
            got = b"\n".join(got.split(b"\n")[2:])
        want = golden_bytes(name, kind)
        if got != want:
            d = util.first_diff(got.decode().splitlines(), want.decode().splitlines())
            raise AssertionError("%s %s differs from the reference\n%s" % (name, kind, d))
        assert len(want) > 0 or kind == "notmatched"


@pytest.mark.parametrize("chunk", ["", "16"])
@pytest.mark.parametrize("name", list(M.VARIANTS))
def test_sinks_write_the_reference_files(name, chunk, tmp_path):
    from usearch12_b200 import build
    cli, replay = build.build_cli(), build.build_format_replay()
    tmp = str(tmp_path)
    q, d = M.write_inputs(name, tmp)
    # the letters the reference's database holds are masked (loaddb.cpp:117-118): -makeudb_usearch is host only
    udb = os.path.join(tmp, "db.udb")
    subprocess.run([cli, "-makeudb_usearch", d, "-output", udb, "-quiet"], check=True)
    hits = os.path.join(tmp, "hits.tsv")
    open(hits, "wb").write(golden_bytes(name, "hits"))
    paths = out_paths(name, tmp)
    cmd = [replay, "-query", q, "-db", udb, "-hits", hits, "-userfields", M.VARIANTS[name][5]] + REPLAY_OPTS[name]
    for k, path in paths.items():
        cmd += [OUT_FLAGS[k], path]
    env = dict(os.environ)
    if chunk:  # several formatting threads per batch, chunks written in input order
        env["USB_FORMAT_CHUNK"] = chunk
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout
    check_outputs(name, paths)


def test_global_segment_fields_are_refused(tmp_path):
    """qseg / tseg / gc read m_HSP.Leni letters from the first M position (alignresult.h:173): for a global
    alignment that runs past the end of the sequence in the reference, so the mirror refuses them."""
    from usearch12_b200 import build
    replay = build.build_format_replay()
    q, d = M.write_inputs("fmt_sz", str(tmp_path))
    hits = os.path.join(str(tmp_path), "hits.tsv")
    open(hits, "wb").write(b"")
    r = subprocess.run([replay, "-query", q, "-db", d, "-hits", hits, "-userout", os.path.join(str(tmp_path), "u"),
                        "-userfields", "query+qseg"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "only supported with -usearch_local" in r.stdout


def test_fastq_queries_and_fastq_outputs(tmp_path):
    """FASTQ query files (fastqseqsource.cpp:8-115: CR LF line ends, '+label' lines, blank lines at the end of the
    file) and -matchedfq / -notmatchedfq (seqdb.cpp:14-28) against the reference binary's files; the hits, -matched
    and -uc are those of the same reads given as FASTA (checked when the fixtures were made)."""
    from usearch12_b200 import build
    replay = build.build_format_replay()
    tmp = str(tmp_path)
    _, d = M.write_inputs("fmt_nt", tmp)
    q = os.path.join(tmp, "q.fq")
    M.write_fastq(q)
    hits = os.path.join(tmp, "hits.tsv")
    open(hits, "wb").write(golden_bytes("fmt_nt", "hits"))
    outs = {k: os.path.join(tmp, "o." + k) for k in ("matchedfq", "notmatchedfq", "matched", "uc")}
    cmd = [replay, "-query", q, "-db", d, "-hits", hits]
    for k, path in outs.items():
        cmd += ["-" + k, path]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for k, path in outs.items():
        assert open(path, "rb").read() == golden_bytes("fmt_fq" if k.endswith("fq") else "fmt_nt", k), k
    # FASTA queries have no qualities (seqdb.cpp:19-20)
    qa, _ = M.write_inputs("fmt_nt", tmp)
    r = subprocess.run([replay, "-query", qa, "-db", d, "-hits", hits, "-matchedfq", outs["matchedfq"]],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Cannot convert FASTA to FASTQ" in r.stdout


@pytest.mark.parametrize("text,msg", [
    ("@a\nACGT\n+\nIII\n", "Bad FASTQ record: 4 bases, 3 quals"),
    ("@a\nAC-T\n+\nIIII\n", "Invalid sequence letter '-'"),
    ("@a\nACGT\n+\nIIII\n\n@b\nAC\n+\nII\n", "Empty line nr 5"),
    ("@a\nACGT\n+\nIIII\nb\nAC\n+\nII\n", "Bad line 5"),
    ("@a\nACGT\n", "Unexpected end-of-file"),
])
def test_fastq_errors_are_the_reference_messages(text, msg, tmp_path):
    from usearch12_b200 import build
    replay = build.build_format_replay()
    q = tmp_path / "bad.fq"
    q.write_text(text)
    d = tmp_path / "db.fa"
    d.write_text(">t\nACGTACGTACGT\n")
    h = tmp_path / "h.tsv"
    h.write_text("")
    r = subprocess.run([replay, "-query", str(q), "-db", str(d), "-hits", str(h)], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and msg in r.stdout, r.stdout


UNIQUES_OLD = {"uniq_sizeout": ["-sizeout"], "uniq_relabel": ["-sizeout", "-relabel", "Uniq", "-minuniquesize", "2"], "uniq_plain": []}


@pytest.mark.parametrize("name", sorted(UNIQUES_OLD) + ["uniq2_" + k for k in sorted(util.UNIQUES2_VARIANTS)])
def test_uniques_writer_matches_reference(name, tmp_path):
    """DerepResult::Write for -fastx_uniques (derepresult.cpp:255-284,689-775,822-844: size order by the reference's
    own quicksort, -sizein sums, -topn, -minuniquesize, -relabel) behind a grouping made on the host by the test tool;
    the product groups on the device (tests/test_gpu_uniques.py)."""
    from usearch12_b200 import build
    replay = build.build_format_replay()
    src = str(tmp_path / "in.fa")
    with gzip.open(os.path.join(util.GOLDEN, "uniq_in.fa.gz"), "rb") as f, open(src, "wb") as g:
        g.write(f.read())
    dst = str(tmp_path / "out.fa")
    extra = UNIQUES_OLD[name] if name in UNIQUES_OLD else util.UNIQUES2_VARIANTS[name[6:]]
    cmd = [replay, "-uniques", src, "-fastaout", dst]
    i = 0
    while i < len(extra):  # the test tool takes "-flag 1" for flags
        if extra[i] in ("-sizein", "-sizeout"):
            cmd += [extra[i]]
            i += 1
        else:
            cmd += extra[i:i + 2]
            i += 2
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    with gzip.open(os.path.join(util.GOLDEN, name + ".fa.gz"), "rb") as f:
        assert open(dst, "rb").read() == f.read()


def _otutab_inputs(tmp):
    for n in ("otutab_reads", "otutab_otus"):
        with gzip.open(os.path.join(util.GOLDEN, n + ".fa.gz"), "rb") as f, open(os.path.join(tmp, n + ".fa"), "wb") as g:
            g.write(f.read())


def check_otutab_outputs(tmp):
    import make_golden_otutab2 as T
    for got, want in (("tab.txt", "otutab.tab.gz"), ("map.txt", "otutab.map.gz"), ("o.biom", "otutab.biom.gz")):
        data = open(os.path.join(tmp, got), "rb").read()
        if got == "o.biom":
            assert b'"date": "' in data and b'"date": ""' not in data
            data = T.blank_date(data)
        with gzip.open(os.path.join(util.GOLDEN, want), "rb") as f:
            assert data == f.read(), want
    # DBHitSink behind -otutab counts one hit per query (dbhitsink.cpp:138-139); -sizeout writes the counts
    sums = json.load(open(os.path.join(util.GOLDEN, "otutab_sha256.json")))
    for k, want in sums.items():
        data = open(os.path.join(tmp, k), "rb").read()
        assert (hashlib.sha256(data).hexdigest(), data.count(b">")) == (want["sha256"], want["seqs"]), k


def test_otutab_sink_writes_the_reference_files(tmp_path):
    """OtuTabSink (otutabsink.cpp:25-76; OTUTable::ToTabbedFile otutab.cpp:247-310, ToJsonFile json.cpp:32-110, the
    -mapout lines) behind the reference's own hit table (tests/golden/otutab.hits.gz): -otutabout, -mapout and
    -biomout byte-identical to the reference binary's files, without a GPU."""
    from usearch12_b200 import build
    replay = build.build_format_replay()
    tmp = str(tmp_path)
    _otutab_inputs(tmp)
    open(os.path.join(tmp, "hits.tsv"), "wb").write(golden_bytes("otutab", "hits"))
    # -dbmatched writes the letters the database holds: masked (loaddb.cpp:117-118)
    subprocess.run([build.build_cli(), "-makeudb_usearch", "otutab_otus.fa", "-output", "otus.udb", "-quiet"], cwd=tmp, check=True)
    r = subprocess.run([replay, "-query", "otutab_reads.fa", "-db", "otus.udb", "-hits", "hits.tsv", "-otutabout",
                        "tab.txt", "-mapout", "map.txt", "-biomout", "o.biom", "-dbmatched", "dbm.fa", "-dbnotmatched", "dbnm.fa",
                        "-notmatched", "nm.fa", "-sizeout"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    check_otutab_outputs(tmp)


def test_userout_needs_userfields(tmp_path):
    """outputsink.cpp:149-156: Die("--userout requires --userfields")."""
    from usearch12_b200 import build
    replay = build.build_format_replay()
    q, d = M.write_inputs("fmt_sz", str(tmp_path))
    hits = os.path.join(str(tmp_path), "hits.tsv")
    open(hits, "wb").write(b"")
    r = subprocess.run([replay, "-query", q, "-db", d, "-hits", hits, "-userout", os.path.join(str(tmp_path), "u")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "--userout requires --userfields" in r.stdout


@pytest.mark.parametrize("name", ["acc_maxhits", "acc_tophit", "acc_tophits"])
def test_hit_selection_matches_reference(name, tmp_path):
    """-maxhits / -top_hit_only / -top_hits_only (HitMgr::GetHitCount / GetHit / GetTopHit, hitmgr.cpp:367-420,466-475):
    SelectHits applied to the reference's unselected hit lists (tests/golden/acc_sel.hits.gz) gives the files the
    reference wrote with the option (tools/make_golden_accept.py)."""
    import make_golden_selection as S
    from usearch12_b200 import build
    replay = build.build_format_replay()
    tmp = str(tmp_path)
    for n in ("acc_q", "acc_db"):
        with gzip.open(os.path.join(util.GOLDEN, n + ".fa.gz"), "rb") as f, open(os.path.join(tmp, n + ".fa"), "wb") as g:
            g.write(f.read())
    open(os.path.join(tmp, "hits.tsv"), "wb").write(golden_bytes("acc_sel", "hits"))
    r = subprocess.run([replay, "-query", "acc_q.fa", "-db", "acc_db.fa", "-hits", "hits.tsv", "-userout", "o.user",
                        "-userfields", S.USERFIELDS, "-uc", "o.uc", "-blast6out", "o.b6"] + S.SELECTIONS[name], cwd=tmp,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for kind in ("user", "uc", "b6"):
        assert open(os.path.join(tmp, "o." + kind), "rb").read() == golden_bytes(name, kind), kind


@pytest.mark.parametrize("args,msg", [
    (["-id", "0.9x", "-strand", "plus"], "Invalid floating-point number '0.9x'"),
    (["-id", "0.9", "-strand", "plus", "-maxaccepts", "abc"], "Invalid integer 'abc'"),
    (["-id", "0.9", "-strand", "plus", "-maxrejects", "-1"], "Invalid integer '-1'"),
])
def test_cli_rejects_malformed_numbers(args, msg):
    """StrToUint / StrToFloat (myutils.cpp:1148-1155,1217-1231) stop the program; no device is needed to get there."""
    from usearch12_b200 import build
    r = subprocess.run([build.build_cli(), "-usearch_global", "x.fa", "-db", "y.fa"] + args, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and msg in r.stdout, r.stdout
