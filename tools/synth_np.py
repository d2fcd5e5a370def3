"""Fast numpy generator of the synthetic usearch_global workloads.

Same statistical shape as tools/gen_synth.py (the pure-python generator the golden fixtures
were made with): DB = n/100 random roots, each target a root mutated at a per-target rate
U(0.03, 0.15) (10 % deletions, 10 % insertions, 80 % substitutions); reads = 250 bp windows of
random targets mutated at U(0, 0.04), 5 % replaced by random sequence.  Everything is seeded.
"""
import numpy as np

LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate_concat(codes, lens, rates, rng):
    """codes: uint8 0..3 concatenated sequences; returns (new codes, new lens)."""
    n = codes.size
    rate_e = np.repeat(rates.astype(np.float32), lens)
    mut = rng.random(n, dtype=np.float32) < rate_e
    k = rng.random(n, dtype=np.float32)
    dele = mut & (k < 0.1)
    ins = mut & (k >= 0.1) & (k < 0.2)
    sub = mut & (k >= 0.2)
    out = codes.copy()
    ns = int(sub.sum())
    out[sub] = (codes[sub] + rng.integers(1, 4, size=ns, dtype=np.uint8)) & 3
    cnt = np.ones(n, dtype=np.uint8)
    cnt[dele] = 0
    cnt[ins] = 2
    new = np.repeat(out, cnt)
    cs = np.cumsum(cnt, dtype=np.int64)
    ins_pos = cs[ins] - 1
    new[ins_pos] = rng.integers(0, 4, size=ins_pos.size, dtype=np.uint8)
    ends = np.cumsum(lens, dtype=np.int64)
    cs0 = np.concatenate([[0], cs])
    new_lens = cs0[ends] - cs0[ends - lens]
    return new, new_lens


def gen_db(n_db, db_len, seed):
    """-> (letters uint8 concatenated, offsets uint64[n_db+1])."""
    rng = np.random.default_rng(seed)
    n_root = max(1, n_db // 100)
    roots = rng.integers(0, 4, size=(n_root, db_len), dtype=np.uint8)
    parts, lens_all = [], []
    chunk = 20000
    for c0 in range(0, n_db, chunk):
        c1 = min(n_db, c0 + chunk)
        idx = np.arange(c0, c1) % n_root
        codes = roots[idx].reshape(-1)
        lens = np.full(c1 - c0, db_len, dtype=np.int64)
        rates = rng.uniform(0.03, 0.15, size=c1 - c0)
        new, nl = _mutate_concat(codes, lens, rates, rng)
        parts.append(LETTERS[new])
        lens_all.append(nl)
    lens_all = np.concatenate(lens_all)
    off = np.zeros(n_db + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens_all)
    return np.concatenate(parts), off


def gen_reads(db, db_off, n_reads, read_len, seed, window=None):
    """-> (letters, offsets uint64[n_reads+1], true target per read (int64, -1 = random read)).
    window=(lo, hi): amplicon variant, every read starts at lo (cluster_fast config)."""
    rng = np.random.default_rng(seed)
    n_db = len(db_off) - 1
    code_of = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code_of[ch] = i
    dlen = (db_off[1:] - db_off[:-1]).astype(np.int64)
    parts, lens_all, truth = [], [], []
    chunk = 100000
    for c0 in range(0, n_reads, chunk):
        m = min(n_reads, c0 + chunk) - c0
        t = rng.integers(0, n_db, size=m)
        span = np.maximum(1, dlen[t] - read_len)
        p = (rng.random(m) * span).astype(np.int64) if window is None else np.full(m, window[0], dtype=np.int64)
        ln = np.minimum(read_len, dlen[t] - p)
        ln = np.maximum(ln, 1)
        start = db_off[t].astype(np.int64) + p
        # ragged gather of the windows
        rel = np.arange(int(ln.sum()), dtype=np.int64) - np.repeat(np.cumsum(ln) - ln, ln)
        codes = code_of[db[np.repeat(start, ln) + rel]]
        rates = rng.uniform(0.0, 0.04, size=m)
        new, nl = _mutate_concat(codes, ln, rates, rng)
        rnd = rng.random(m) < 0.05
        # random reads: overwrite the read's letters with noise
        e = np.cumsum(nl)
        b = e - nl
        for i in np.nonzero(rnd)[0]:
            new[b[i]:e[i]] = rng.integers(0, 4, size=int(nl[i]), dtype=np.uint8)
        tt = t.copy()
        tt[rnd] = -1
        parts.append(LETTERS[new])
        lens_all.append(nl)
        truth.append(tt)
    lens_all = np.concatenate(lens_all)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens_all)
    return np.concatenate(parts), off, np.concatenate(truth)


def write_fasta(path, letters, off, prefix, lo=0, hi=None):
    hi = len(off) - 1 if hi is None else hi
    buf = letters.tobytes()
    with open(path, "wb") as f:
        for i in range(lo, hi):
            f.write(b">%s%d\n" % (prefix.encode(), i))
            f.write(buf[int(off[i]):int(off[i + 1])])
            f.write(b"\n")


# ---------------------------------------------------------------- proteins (config 5)
AA_LETTERS = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def _mutate_concat_aa(codes, lens, rates, rng):
    """aa mutation mix: 5 % deletion, 5 % insertion of a random residue after, 90 % substitution by
    a uniformly random residue."""
    n = codes.size
    rate_e = np.repeat(rates.astype(np.float32), lens)
    mut = rng.random(n, dtype=np.float32) < rate_e
    k = rng.random(n, dtype=np.float32)
    dele = mut & (k < 0.05)
    ins = mut & (k >= 0.05) & (k < 0.10)
    sub = mut & (k >= 0.10)
    out = codes.copy()
    out[sub] = rng.integers(0, 20, size=int(sub.sum()), dtype=np.uint8)
    cnt = np.ones(n, dtype=np.uint8)
    cnt[dele] = 0
    cnt[ins] = 2
    new = np.repeat(out, cnt)
    cs = np.cumsum(cnt, dtype=np.int64)
    ins_pos = cs[ins] - 1
    new[ins_pos] = rng.integers(0, 20, size=ins_pos.size, dtype=np.uint8)
    ends = np.cumsum(lens, dtype=np.int64)
    cs0 = np.concatenate([[0], cs])
    new_lens = cs0[ends] - cs0[ends - lens]
    return new, new_lens


def gen_aa(n_db, length, n_q, seed):
    """-> (db letters, db offsets, query letters, query offsets, true target per query).
    DB: n_db/20 random roots, target = root mutated at U(0.05, 0.5); query = random target mutated
    at U(0.05, 0.4)."""
    rng = np.random.default_rng(seed)
    n_root = max(1, n_db // 20)
    roots = rng.integers(0, 20, size=(n_root, length), dtype=np.uint8)
    idx = np.arange(n_db) % n_root
    dcodes, dlens = _mutate_concat_aa(roots[idx].reshape(-1), np.full(n_db, length, dtype=np.int64),
                                      rng.uniform(0.05, 0.5, size=n_db), rng)
    doff = np.zeros(n_db + 1, dtype=np.uint64)
    doff[1:] = np.cumsum(dlens)
    qparts, qlens, truth = [], [], []
    chunk = 50000
    for c0 in range(0, n_q, chunk):
        m = min(n_q, c0 + chunk) - c0
        t = rng.integers(0, n_db, size=m)
        ln = dlens[t]
        start = doff[t].astype(np.int64)
        rel = np.arange(int(ln.sum()), dtype=np.int64) - np.repeat(np.cumsum(ln) - ln, ln)
        codes = dcodes[np.repeat(start, ln) + rel]
        new, nl = _mutate_concat_aa(codes, ln, rng.uniform(0.05, 0.4, size=m), rng)
        qparts.append(new)
        qlens.append(nl)
        truth.append(t)
    qlens = np.concatenate(qlens)
    qoff = np.zeros(n_q + 1, dtype=np.uint64)
    qoff[1:] = np.cumsum(qlens)
    return AA_LETTERS[dcodes], doff, AA_LETTERS[np.concatenate(qparts)], qoff, np.concatenate(truth)
