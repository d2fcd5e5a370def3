""".udb files: usb_udb_write must produce, byte for byte, the file the
reference's -makeudb_usearch writes (udbio.cpp:281-364, seqdbio.cpp:17-135), and usb_udb_read must
take such a file apart again.  Host-only entry points: no GPU needed."""
import ctypes as C
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import util
from usearch12_b200 import capi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "usearch12")
DIGESTS = json.load(open(os.path.join(GOLDEN, "udb_sha256.json")))


def params_for(name):
    p = capi.default_params()
    if "aa" in name:
        capi.lib().usb_set_local(C.byref(p), 0, 1e-5)
    return p


@pytest.mark.parametrize("name", sorted(DIGESTS))
def test_udb_write_is_byte_identical_to_the_reference(name, tmp_path):
    """Keys "<file>:w<n>" are written with -wordlength n (udbparams.cpp:58-81)."""
    fname, _, wl = name.partition(":w")
    labels, seqs = util.read_fasta(os.path.join(GOLDEN, fname))
    path = str(tmp_path / "ours.udb")
    p = params_for(fname)
    if wl:
        p.word_length = int(wl)
    capi.udb_write(path, labels, seqs, p)
    ours = open(path, "rb").read()
    assert len(ours) == DIGESTS[name]["bytes"]
    assert hashlib.sha256(ours).hexdigest() == DIGESTS[name]["sha256"]
    assert capi.lib().usb_udb_probe(path.encode()) == 1
    fa = str(tmp_path / "in.fa")
    with gzip.open(os.path.join(GOLDEN, fname), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    extra = ["-wordlength", wl] if wl else []
    if wl:  # the command line of the host driver writes the same file
        from usearch12_b200 import build
        cli_out = str(tmp_path / "cli.udb")
        subprocess.run([build.build_cli(), "-makeudb_usearch", fa, "-output", cli_out, "-quiet"] + extra, check=True)
        assert open(cli_out, "rb").read() == ours
    if os.path.exists(REF):  # this container: the reference binary itself, byte by byte
        ref = str(tmp_path / "ref.udb")
        subprocess.run([REF, "-makeudb_usearch", fa, "-output", ref, "-quiet"] + extra, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        assert open(ref, "rb").read() == ours


def test_udb_read_round_trip(tmp_path):
    labels, seqs = util.read_fasta(os.path.join(GOLDEN, "db.fa.gz"))
    path = str(tmp_path / "db.udb")
    capi.udb_write(path, labels, seqs)
    u = capi.Udb(path)
    assert u.n_seq == len(seqs) and u.is_nucleo and u.word_length == 8
    assert u.labels == labels
    # stored letters = the masked database: same letters, lower case where FastMaskSeq masked
    assert [s.upper() for s in u.seqs] == [s.upper().encode() for s in seqs]
    assert any(s != s.upper() for s in u.seqs)
    # rows: ascending, every target at most once, and exactly the targets that contain the word unmasked
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
    want = {}
    for t, s in enumerate(u.seqs):
        for i in range(len(s) - 7):
            w = 0
            for c in s[i:i + 8]:
                if c not in code:
                    w = -1
                    break
                w = w * 4 + code[c]
            if w >= 0:
                want.setdefault(w, set()).add(t)
    rng = np.random.default_rng(5)
    words = list(rng.integers(0, 65536, 300)) + list(want)[:300]
    for w in words:
        row = u.row(int(w))
        assert list(row) == sorted(want.get(int(w), ()))
    u.close()


def test_udb_read_rejects_other_files(tmp_path):
    bad = tmp_path / "x.udb"
    bad.write_bytes(b">seq\nACGT\n")
    assert capi.lib().usb_udb_probe(str(bad).encode()) == 0
    h = C.c_void_p()
    assert capi.lib().usb_udb_read(str(bad).encode(), C.byref(h)) == -1
    assert b"not a .udb file" in capi.lib().usb_last_error()
    labels, seqs = util.read_fasta(os.path.join(GOLDEN, "db.fa.gz"))
    path = str(tmp_path / "db.udb")
    capi.udb_write(path, labels[:50], seqs[:50])
    whole = open(path, "rb").read()
    trunc = tmp_path / "t.udb"
    trunc.write_bytes(whole[:len(whole) - 1000])
    assert capi.lib().usb_udb_read(str(trunc).encode(), C.byref(h)) == -1
    assert b"truncated" in capi.lib().usb_last_error()
    assert capi.lib().usb_udb_write(path.encode(), C.byref(capi.default_params()), None, None, None, 0) == -1


def test_udb_read_rejects_hostile_sizes(tmp_path):
    """A .udb whose size fields exceed the file must come back as an error code, not as an
    exception or abort through the C ABI (sizes are checked against the bytes left in the file)."""
    import random
    labels = ["t%d" % i for i in range(20)]
    rng = random.Random(3)
    seqs = ["".join(rng.choice("ACGT") for _ in range(300)) for _ in labels]
    path = str(tmp_path / "ok.udb")
    capi.udb_write(path, labels, seqs)
    raw = bytearray(open(path, "rb").read())
    u = capi.Udb(path)
    assert u.n_seq == 20
    u.close()
    # header is 200 bytes, then sizes[65536] (u32): blow every row size up
    bad = bytearray(raw)
    bad[200:200 + 4 * 65536] = b"\xff" * (4 * 65536)
    p2 = str(tmp_path / "bad.udb")
    open(p2, "wb").write(bad)
    h = C.c_void_p()
    rc = capi.lib().usb_udb_read(p2.encode(), C.byref(h))
    assert rc != 0 and not h.value
    assert b"inconsistent" in capi.lib().usb_last_error() or b"truncated" in capi.lib().usb_last_error()
    # truncated file
    p3 = str(tmp_path / "short.udb")
    open(p3, "wb").write(raw[: len(raw) // 2])
    rc = capi.lib().usb_udb_read(p3.encode(), C.byref(h))
    assert rc != 0


def test_parallel_fasta_parse_equals_serial(tmp_path):
    """SeqDB::FromFasta cuts large files at '>' lines and parses the pieces in threads: the .udb written
    from the result and the warnings must not depend on the number of pieces."""
    import random
    import subprocess
    from usearch12_b200 import build
    rng = random.Random(12)
    lines = []
    for i in range(3000):
        L = rng.choice([0, 1, 7, 8, 60, 61, 200, 900]) if i % 50 == 7 else rng.randrange(30, 400)
        s = "".join(rng.choice("ACGT") for _ in range(L))
        if i % 11 == 0:
            s = s.lower()
        if i % 13 == 0 and L > 10:
            s = s[:5] + "N-.*1" + s[5:]
        lines.append(">seq%d some text; with >inside\r" % i if i % 17 == 0 else ">seq%d" % i)
        w = rng.choice([60, 80, 10 ** 6])
        for k in range(0, len(s), w):
            lines.append(s[k:k + w] + ("\r" if i % 17 == 0 else ""))
        if i % 29 == 0:
            lines.append("")
    fa = tmp_path / "in.fa"
    fa.write_text("\n".join(lines) + ("\n" if rng.random() < 0.5 else ""))
    cli = build.build_cli()
    outs = []
    for t in ("1", "2", "7", "33"):
        udb = tmp_path / ("o%s.udb" % t)
        r = subprocess.run([cli, "-makeudb_usearch", str(fa), "-output", str(udb)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, env=dict(os.environ, USB_FASTA_THREADS=t))
        assert r.returncode == 0, r.stdout
        outs.append((udb.read_bytes(), [l for l in r.stdout.splitlines() if "WARNING" in l.upper() or "Empty sequence" in l]))
    assert all(o == outs[0] for o in outs[1:])
    assert len(outs[0][0]) > 100000 and any("Empty sequence" in l for l in outs[0][1])
