"""ctypes binding of the CPU ORACLE (oracle/_build/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product package (usearch12_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


class Params(C.Structure):
    _fields_ = [
        ("is_nucleo", C.c_int), ("id", C.c_float), ("maxaccepts", C.c_uint), ("maxrejects", C.c_uint),
        ("strand_both", C.c_int), ("word_length", C.c_uint), ("big", C.c_uint), ("bump", C.c_uint),
        ("stepwords", C.c_uint), ("band", C.c_uint), ("minhsp", C.c_uint), ("hspw", C.c_uint),
        ("xdrop_nw", C.c_float), ("match", C.c_float), ("mismatch", C.c_float), ("dbmask_fast", C.c_int),
        ("cluster_mode", C.c_int), ("fulldp", C.c_int), ("local", C.c_int), ("evalue", C.c_float),
        ("xdrop_u", C.c_float), ("xdrop_g", C.c_float), ("lopen", C.c_float), ("lext", C.c_float),
        ("ka_dbsize", C.c_float),
    ]


class Hit(C.Structure):
    _fields_ = [
        ("query", C.c_uint32), ("target", C.c_uint32), ("strand", C.c_uint8),
        ("ids", C.c_uint32), ("mism", C.c_uint32), ("intgaps", C.c_uint32), ("opens", C.c_uint32),
        ("first_mq", C.c_uint32), ("first_mt", C.c_uint32), ("last_mq", C.c_uint32), ("last_mt", C.c_uint32),
        ("first_mcol", C.c_uint32), ("alnlen", C.c_uint32), ("ql", C.c_uint32), ("tl", C.c_uint32),
        ("loi", C.c_uint32), ("loj", C.c_uint32), ("leni", C.c_uint32), ("lenj", C.c_uint32),
        ("raw", C.c_double), ("evalue", C.c_double), ("bits", C.c_double),
        ("path", C.c_char_p),
    ]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "uso.c")):
            build()
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.uso_default_params.argtypes = [C.POINTER(Params), C.c_int]
        L.uso_set_amino.argtypes = [C.POINTER(Params)]
        L.uso_xdrop_fwd.argtypes = [C.POINTER(Params), C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_float,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_char_p]
        L.uso_xdrop_fwd.restype = C.c_float
        L.uso_local_align_pos.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                          vp, C.POINTER(C.c_float), C.c_char_p]
        L.uso_local_align_pos.restype = C.c_int
        L.uso_db_create.argtypes = [C.POINTER(Params)]
        L.uso_db_create.restype = vp
        L.uso_db_add.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p]
        L.uso_db_add.restype = C.c_uint32
        L.uso_db_free.argtypes = [vp]
        L.uso_db_seq_count.argtypes = [vp]
        L.uso_db_seq_count.restype = C.c_uint32
        L.uso_db_seq.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
        L.uso_db_seq.restype = C.POINTER(C.c_uint8)
        L.uso_db_row.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
        L.uso_db_row.restype = C.POINTER(C.c_uint32)
        L.uso_searcher_create.argtypes = [vp, C.POINTER(Params)]
        L.uso_searcher_create.restype = vp
        L.uso_searcher_free.argtypes = [vp]
        L.uso_search.argtypes = [vp, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(C.POINTER(Hit)),
                                 C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.uso_search.restype = C.c_uint
        L.uso_hits_free.argtypes = [C.POINTER(Hit), C.c_uint]
        L.uso_rank_candidates.argtypes = [vp, C.c_char_p, C.c_uint32, vp, vp, vp]
        L.uso_rank_candidates.restype = C.c_uint
        L.uso_rank_candidates_big.argtypes = [vp, C.c_char_p, C.c_uint32, vp, vp]
        L.uso_rank_candidates_big.restype = C.c_uint
        L.uso_global_hsps.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, vp, C.POINTER(C.c_uint), vp,
                                      C.c_uint, C.POINTER(C.c_float)]
        L.uso_global_hsps.restype = C.c_uint
        L.uso_viterbi_band.argtypes = [C.POINTER(Params), C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.uso_viterbi_band.restype = C.c_float
        L.uso_global_align.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p]
        L.uso_global_align.restype = C.c_int
        L.uso_fastmask.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.uso_revcomp.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.uso_compress_path.argtypes = [C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


def default_params(cluster_fast=False, amino=False, **kw):
    p = Params()
    lib().uso_default_params(C.byref(p), int(cluster_fast))
    if amino:
        lib().uso_set_amino(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _b(s):
    return s.encode() if isinstance(s, str) else bytes(s)


def fastmask(seq):
    s = _b(seq)
    out = C.create_string_buffer(len(s) + 1)
    lib().uso_fastmask(s, len(s), out)
    return out.raw[:len(s)]


def revcomp(seq):
    s = _b(seq)
    out = C.create_string_buffer(len(s) + 1)
    lib().uso_revcomp(s, len(s), out)
    return out.raw[:len(s)]


def compress_path(path):
    out = C.create_string_buffer(2 * len(path) + 16)
    lib().uso_compress_path(_b(path), out)
    return out.value.decode()


def viterbi_band(params, a, b, left_a, left_b, right_a, right_b):
    a, b = _b(a), _b(b)
    out = C.create_string_buffer(len(a) + len(b) + 2)
    sc = lib().uso_viterbi_band(C.byref(params), a, len(a), b, len(b), int(left_a), int(left_b), int(right_a),
                                int(right_b), out)
    return out.value.decode(), sc


def xdrop_fwd(params, a, b, x):
    """XDropFwdFastMem alone -> (score, leni, lenj, path)."""
    a, b = _b(a), _b(b)
    out = C.create_string_buffer(len(a) + len(b) + 2)
    li, lj = C.c_uint32(), C.c_uint32()
    sc = lib().uso_xdrop_fwd(C.byref(params), a, len(a), b, len(b), float(x), C.byref(li), C.byref(lj), out)
    return sc, li.value, lj.value, out.value.decode()


class DB:
    def __init__(self, seqs, params, labels=None):
        self.params = params
        self.h = lib().uso_db_create(C.byref(params))
        for i, s in enumerate(seqs):
            s = _b(s)
            lib().uso_db_add(self.h, s, len(s), _b(labels[i]) if labels else b"")
        self.n = len(seqs)

    def seq(self, i):
        n = C.c_uint32()
        p = lib().uso_db_seq(self.h, i, C.byref(n))
        return bytes(np.ctypeslib.as_array(p, shape=(n.value,))) if n.value else b""

    def row(self, word):
        n = C.c_uint32()
        p = lib().uso_db_row(self.h, word, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def __del__(self):
        try:
            lib().uso_db_free(self.h)
        except Exception:
            pass


class Searcher:
    def __init__(self, db, params=None):
        self.db = db
        self.params = params or db.params
        self.h = lib().uso_searcher_create(db.h, C.byref(self.params))

    def search(self, seq, qindex=0):
        """-> list of dicts in the reference's output order."""
        s = _b(seq)
        hits = C.POINTER(Hit)()
        n = C.c_uint(0)
        cap = C.c_uint(0)
        k = lib().uso_search(self.h, qindex, s, len(s), C.byref(hits), C.byref(n), C.byref(cap))
        out = []
        for i in range(k):
            h = hits[i]
            d = {f[0]: getattr(h, f[0]) for f in Hit._fields_ if f[0] != "path"}
            d["path"] = h.path.decode()
            out.append(d)
        if n.value:
            lib().uso_hits_free(hits, n.value)
        return out

    def rank(self, seq):
        s = _b(seq)
        N = self.db.n
        U = np.zeros(max(N, 1), np.uint32)
        ct = np.zeros(max(N, 1), np.uint32)
        cu = np.zeros(max(N, 1), np.uint32)
        k = lib().uso_rank_candidates(self.h, s, len(s), U.ctypes.data, ct.ctypes.data, cu.ctypes.data)
        return U[:N], ct[:k], cu[:k]

    def rank_big(self, seq):
        """Candidate order of the big-database path -> (targets, U)."""
        s = _b(seq)
        N = self.db.n
        ct = np.zeros(max(N, 1), np.uint32)
        cu = np.zeros(max(N, 1), np.uint32)
        k = lib().uso_rank_candidates_big(self.h, s, len(s), ct.ctypes.data, cu.ctypes.data)
        return ct[:k], cu[:k]

    def global_hsps(self, q, t, max_hsp=256):
        q, t = _b(q), _b(t)
        ung = np.zeros(4 * max_hsp, np.uint32)
        ch = np.zeros(4 * max_hsp, np.uint32)
        nu = C.c_uint()
        fid = C.c_float()
        nc = lib().uso_global_hsps(self.h, q, len(q), t, len(t), ung.ctypes.data, C.byref(nu), ch.ctypes.data, max_hsp,
                                   C.byref(fid))
        return ung[:4 * min(nu.value, max_hsp)].reshape(-1, 4), ch[:4 * min(nc, max_hsp)].reshape(-1, 4), fid.value

    def local_align_pos(self, q, t, qpos, tpos):
        """LocalAligner::AlignPos -> None or (loi, loj, leni, lenj, score, path)."""
        q, t = _b(q), _b(t)
        out = C.create_string_buffer(len(q) + len(t) + 8)
        h4 = np.zeros(4, np.uint32)
        sc = C.c_float()
        ok = lib().uso_local_align_pos(self.h, q, len(q), t, len(t), qpos, tpos, h4.ctypes.data, C.byref(sc), out)
        return (int(h4[0]), int(h4[1]), int(h4[2]), int(h4[3]), sc.value, out.value.decode()) if ok else None

    def global_align(self, q, t):
        q, t = _b(q), _b(t)
        out = C.create_string_buffer(len(q) + len(t) + 2)
        ok = lib().uso_global_align(self.h, q, len(q), t, len(t), out)
        return out.value.decode() if ok else None

    def __del__(self):
        try:
            lib().uso_searcher_free(self.h)
        except Exception:
            pass
