"""GPU parity of the cluster_fast greedy centroid loop (usb_cluster_round) against the oracle's
sequential loop (oracle/uso.c, pinned against the reference binary's -cluster_fast output by
tools/pin_oracle.sh-style runs: byte-identical .uc and centroids on 25 715 reads, three -sort modes)."""
import random

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def make_reads(seed, n_random, n_amp, n_db=300, n_amp_db=40):
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    from gen_synth import generate
    rng = random.Random(seed)
    db, reads = generate(ndb=n_db, dblen=1000, nq=n_random, qlen=250, seed=seed, nroot=6)
    out = [r[1] for r in reads]
    for k in range(n_amp):
        t = rng.randrange(n_amp_db)
        L = rng.choice([250, 250, 240, 200, 251])
        out.append(util.mutate(db[t][300:300 + L], rng.uniform(0, 0.03), rng))
    out += ["ACGT", "N" * 60, out[3].lower(), "A" * 250]
    rng.shuffle(out)
    return out


def oracle_cluster(reads, **kw):
    from oracle import uso_py as O
    p = O.default_params(cluster_fast=True, **kw)
    db = O.DB([], p)
    s = O.Searcher(db, p)
    assign, paths = [], []
    import ctypes as C
    for i, r in enumerate(reads):
        hits = s.search(r, i)
        if hits:
            assign.append(hits[0]["target"])
            paths.append((hits[0]["ids"], hits[0]["alnlen"], hits[0]["path"]))
        else:
            b = r.encode()
            assign.append(O.lib().uso_db_add(db.h, b, len(b), b""))
            db.n += 1
            paths.append(None)
    return assign, paths


def gpu_cluster(reads, block, **kw):
    from usearch12_b200 import capi
    p = capi.default_params(cluster_fast=True, **kw)
    ix = capi.Index([], p)
    s = capi.Searcher(ix, p)
    data, off = capi.pack_seqs(reads)
    assign, paths, rounds = [], [], 0
    pos = 0
    while pos < len(reads):
        n = min(block, len(reads) - pos)
        sub_off = off[pos:pos + n + 1]
        ncom, cidx, res = s.cluster_round(data, sub_off)
        assert ncom >= 1
        for q in range(ncom):
            assign.append(int(cidx[q]))
            b, e = int(res.qoff[q]), int(res.qoff[q + 1])
            if e > b:
                h = res.hits[b]
                paths.append((int(h["ids"]), int(h["alnlen"]), res.path(h)))
            else:
                paths.append(None)
        pos += ncom
        rounds += 1
    return assign, paths, rounds


@pytest.mark.parametrize("block", [1, 64, 4096])
def test_cluster_rounds_equal_sequential_loop(block):
    reads = make_reads(7, 500, 700)
    want_a, want_p = oracle_cluster(reads)
    got_a, got_p, rounds = gpu_cluster(reads, block)
    assert got_a == want_a
    assert got_p == want_p
    assert len(set(want_a)) > 50
    if block > 1:
        assert rounds < len(reads)


def test_cluster_crosses_big_threshold():
    """-big 150: the database switches to the UDBSearchBig path in mid-run, like a 1M-read
    cluster_fast crossing 100 000 centroids."""
    reads = make_reads(11, 600, 300)
    want_a, want_p = oracle_cluster(reads, big=150)
    got_a, got_p, _ = gpu_cluster(reads, 512, big=150)
    assert got_a == want_a
    assert got_p == want_p
    assert len(set(want_a)) > 200


@pytest.mark.parametrize("sort", ["none", "length", "size"])
def test_cluster_fast_cli_byte_identical_to_reference(sort, tmp_path):
    """C++ driver (-cluster_fast): .uc and centroids FASTA byte-identical to the reference binary's
    (tests/golden/cluster_*.gz, made with oracle/_ref/usearch12 -cluster_fast -id 0.97 -threads 1)."""
    import gzip
    import os
    import subprocess
    from usearch12_b200 import build
    cli = build.build_cli()
    reads = os.path.join(str(tmp_path), "g.fa")
    with gzip.open(os.path.join(util.GOLDEN, "cluster_reads.fa.gz"), "rb") as fi, open(reads, "wb") as fo:
        fo.write(fi.read())
    uc, cen = os.path.join(str(tmp_path), "o.uc"), os.path.join(str(tmp_path), "o.fa")
    cmd = [cli, "-cluster_fast", reads, "-id", "0.97", "-uc", uc, "-centroids", cen, "-quiet"]
    if sort != "none":
        cmd += ["-sort", sort]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for got, name in ((uc, "cluster_%s.uc.gz" % sort), (cen, "cluster_%s.centroids.fa.gz" % sort)):
        with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
            want = f.read().splitlines()
        d = util.first_diff(open(got).read().splitlines(), want)
        assert d is None, "%s\n%s" % (name, d)


@pytest.mark.parametrize("variant", sorted(util.CLUSTER_SIZE_VARIANTS))
def test_cluster_fast_size_annotations_byte_identical_to_reference(variant, tmp_path):
    """-sizein / -sizeout / -relabel / -minsize (clustersink.cpp:118-143,217-272, derepresult.cpp:211-225,
    clusterfast.cpp:38-79): .uc and centroids identical to the reference binary's
    (tools/make_golden_cluster_sizes.py)."""
    import gzip
    import os
    import subprocess
    from usearch12_b200 import build
    cli = build.build_cli()
    reads = os.path.join(str(tmp_path), "r.fa")
    util.sized_cluster_reads(reads)
    uc, cen = os.path.join(str(tmp_path), "o.uc"), os.path.join(str(tmp_path), "o.fa")
    cmd = [cli, "-cluster_fast", reads, "-id", "0.97", "-uc", uc, "-centroids", cen, "-quiet"] + util.CLUSTER_SIZE_VARIANTS[variant]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for got, name in ((uc, "cluster_%s.uc.gz" % variant), (cen, "cluster_%s.centroids.fa.gz" % variant)):
        with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
            want = f.read().splitlines()
        d = util.first_diff(open(got).read().splitlines(), want)
        assert d is None, "%s\n%s" % (name, d)


@pytest.mark.parametrize("variant", sorted(util.SMALLMEM_VARIANTS))
def test_cluster_smallmem_byte_identical_to_reference(variant, tmp_path):
    """-cluster_smallmem (clustersmallmem.cpp:50-143): no dereplication, input order checked against
    -sortedby, Terminator 1/32 (terminator.cpp:22-31); .uc and centroids identical to the reference
    binary's (tools/make_golden_cluster_sizes.py)."""
    import gzip
    import os
    import subprocess
    from usearch12_b200 import build
    cli = build.build_cli()
    order, extra = util.SMALLMEM_VARIANTS[variant]
    reads = os.path.join(str(tmp_path), "r.fa")
    util.smallmem_reads(reads, order)
    uc, cen = os.path.join(str(tmp_path), "o.uc"), os.path.join(str(tmp_path), "o.fa")
    r = subprocess.run([cli, "-cluster_smallmem", reads, "-id", "0.97", "-uc", uc, "-centroids", cen, "-quiet"] + extra,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    for got, name in ((uc, "cluster_%s.uc.gz" % variant), (cen, "cluster_%s.centroids.fa.gz" % variant)):
        with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
            want = f.read().splitlines()
        d = util.first_diff(open(got).read().splitlines(), want)
        assert d is None, "%s\n%s" % (name, d)


def test_cluster_smallmem_refuses_unsorted_input(tmp_path):
    import os
    import subprocess
    from usearch12_b200 import build
    reads = os.path.join(str(tmp_path), "r.fa")
    util.smallmem_reads(reads, "none")
    r = subprocess.run([build.build_cli(), "-cluster_smallmem", reads, "-id", "0.97", "-uc", os.path.join(str(tmp_path), "o.uc")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0 and "Not sorted by length" in r.stdout


@pytest.mark.parametrize("kind,n_default", [("window", 120000), ("amplicon", 200000)])
def test_cluster_fast_120k_reads_identical_to_reference_across_the_big_switch(kind, n_default, tmp_path):
    """BASELINE config 3 on real volume: the first 120 000 window-random (200 000 amplicon) reads of the
    bench workload give more than 100 000 clusters, so the run crosses -big (100 000 targets:
    udbusortedsearcher.cpp:39-58) and the last rounds take the big-database path (sampled words,
    first-touch order, k_rank_big's many-survivor selection).  .uc byte-identical to the unmodified
    reference binary (-threads 1)."""
    import os
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tools"))
    import synth_np
    from usearch12_b200 import build
    ref = os.path.join(util.ROOT, "oracle", "_ref", "usearch12")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/usearch12 not built")
    n = int(os.environ.get("USB_CLUSTER_TEST_READS", str(n_default)))
    db, db_off = synth_np.gen_db(100000, 1500, seed=4)
    reads, r_off, _ = synth_np.gen_reads(db, db_off, n, 250, seed=3000, window=(500, 750) if kind == "amplicon" else None)
    fa = str(tmp_path / "r.fa")
    synth_np.write_fasta(fa, reads, r_off, "r")
    out = {}
    for name, exe, extra in (("ref", ref, ["-threads", "1"]), ("usb", build.build_cli(), [])):
        uc = str(tmp_path / (name + ".uc"))
        subprocess.run([exe, "-cluster_fast", fa, "-id", "0.97", "-uc", uc, "-quiet"] + extra, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1200)
        out[name] = open(uc, "rb").read()
    n_clusters = sum(1 for l in out["ref"].splitlines() if l.startswith(b"C\t"))
    assert n_clusters > 100000, n_clusters
    assert out["usb"] == out["ref"]
