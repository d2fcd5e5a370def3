"""ctypes binding of libusb200.so (include/usb200.h).

This is the stub a Python caller of the reference's path would add; it holds no algorithm.
The library is CUDA-only: loading works without a GPU (symbol checks), compute calls raise
UsbError when no device is present -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))


class UsbError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("is_nucleo", C.c_int32), ("id", C.c_float),
        ("maxaccepts", C.c_uint32), ("maxrejects", C.c_uint32), ("strand_both", C.c_int32),
        ("word_length", C.c_uint32), ("big", C.c_uint32), ("bump", C.c_uint32), ("stepwords", C.c_uint32),
        ("band", C.c_uint32), ("minhsp", C.c_uint32), ("hspw", C.c_uint32), ("xdrop_nw", C.c_float),
        ("match", C.c_float), ("mismatch", C.c_float), ("gap_open", C.c_float), ("gap_ext", C.c_float),
        ("term_gap_open", C.c_float), ("term_gap_ext", C.c_float), ("dbmask", C.c_int32),
        ("cluster_mode", C.c_int32), ("fulldp", C.c_int32), ("local", C.c_int32), ("evalue", C.c_float),
        ("xdrop_u", C.c_float), ("xdrop_g", C.c_float), ("lopen", C.c_float), ("lext", C.c_float),
        ("ka_dbsize", C.c_float),
        ("accept_flags", C.c_uint32), ("maxid", C.c_float), ("mincols", C.c_uint32), ("maxgaps", C.c_uint32),
        ("maxdiffs", C.c_uint32), ("mindiffs", C.c_uint32), ("query_cov", C.c_float), ("max_query_cov", C.c_float),
        ("target_cov", C.c_float), ("max_target_cov", C.c_float), ("abskew", C.c_float), ("min_sizeratio", C.c_float),
        ("minqt", C.c_float), ("maxqt", C.c_float), ("minsl", C.c_float), ("maxsl", C.c_float),
        ("termid", C.c_float), ("termidd", C.c_float),
    ]


# usb_params.accept_flags (include/usb200.h USB_ACC_*)
ACC = dict(self=0x1, notself=0x2, selfid=0x4, maxid=0x8, mincols=0x10, maxgaps=0x20, query_cov=0x40, max_query_cov=0x80,
           target_cov=0x100, max_target_cov=0x200, maxdiffs=0x400, mindiffs=0x800, abskew=0x1000, min_sizeratio=0x2000,
           minqt=0x4000, maxqt=0x8000, minsl=0x10000, maxsl=0x20000, termid=0x40000, termidd=0x80000)


HIT_DTYPE = np.dtype([
    ("query", "<u4"), ("target", "<u4"), ("strand", "<u4"), ("rank", "<u4"),
    ("ids", "<u4"), ("mism", "<u4"), ("intgaps", "<u4"), ("opens", "<u4"),
    ("first_mq", "<u4"), ("first_mt", "<u4"), ("last_mq", "<u4"), ("last_mt", "<u4"),
    ("first_mcol", "<u4"), ("alnlen", "<u4"), ("ql", "<u4"), ("tl", "<u4"),
    ("run_off", "<u4"), ("run_cnt", "<u4"), ("raw", "<i4"), ("sub", "<u4"),
])
QSTAT_DTYPE = np.dtype([
    ("n_cand", "<u4"), ("n_tried", "<u4"), ("n_hspfail", "<u4"), ("n_dp", "<u4"), ("dp_cells", "<u4"),
    ("n_accept", "<u4"), ("seq_bytes", "<u4"),
])

# every symbol include/usb200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "usb_default_params", "usb_last_error", "usb_device_count", "usb_index_create", "usb_index_append", "usb_index_reserve",
    "usb_index_free",
    "usb_index_seq_count", "usb_index_posting_count", "usb_index_posting_width", "usb_index_row", "usb_index_seq",
    "usb_searcher_create", "usb_searcher_free", "usb_search_batch", "usb_batch_upload", "usb_batch_run",
    "usb_batch_download", "usb_cluster_round", "usb_batch_counters", "usb_batch_kernel_ms", "usb_index_set_attrs", "usb_batch_set_query_attrs",
    "usb_searcher_launch_count", "usb_batch_export_hits_device",
    "usb_result_hit_count", "usb_result_hits", "usb_result_runs", "usb_result_query_offsets",
    "usb_result_qstats", "usb_result_free", "usb_result_path", "usb_rank_batch", "usb_align_pairs",
    "usb_viterbi_batch", "usb_set_local", "usb_set_amino", "usb_local_evalue", "usb_params_evalue", "usb_local_pairs",
    "usb_udb_write", "usb_udb_probe", "usb_udb_read", "usb_udb_free", "usb_udb_seq_count", "usb_udb_is_nucleo",
    "usb_udb_word_length", "usb_udb_seqs", "usb_udb_label", "usb_udb_row", "usb_debug_half_row", "usb_derep_full",
]

_lib = None


def derep_full(seqs, device=0):
    """DerepFull (derepfull.cpp:130-212) on the device -> (uniq_of[n], n_uniq)."""
    data, off = pack_seqs(seqs)
    out = np.zeros(max(1, len(seqs)), np.uint32)
    nu = C.c_uint32(0)
    check(lib().usb_derep_full(device, _ptr(data), _ptr(off), len(seqs), _ptr(out), C.byref(nu)))
    return out[:len(seqs)], nu.value


def lib():
    """Loads (building if stale and nvcc is present) libusb200.so."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = C.CDLL(path)
    vp, u8p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.usb_last_error.restype = C.c_char_p
    L.usb_default_params.argtypes = [C.POINTER(Params), C.c_int]
    L.usb_default_params.restype = None
    L.usb_index_create.argtypes = [C.c_int, C.POINTER(Params), vp, vp, C.c_uint32, C.POINTER(vp)]
    L.usb_index_append.argtypes = [vp, vp, vp, C.c_uint32]
    L.usb_index_reserve.argtypes = [vp, C.c_uint32, C.c_uint64]
    L.usb_index_free.argtypes = [vp]
    L.usb_index_free.restype = None
    L.usb_index_seq_count.argtypes = [vp]
    L.usb_index_seq_count.restype = C.c_uint32
    L.usb_index_posting_count.argtypes = [vp]
    L.usb_index_posting_count.restype = C.c_uint64
    L.usb_index_posting_width.argtypes = [vp]
    L.usb_index_posting_width.restype = C.c_uint32
    L.usb_derep_full.argtypes = [C.c_int, vp, vp, C.c_uint32, vp, u32p]
    L.usb_udb_write.argtypes = [C.c_char_p, C.POINTER(Params), vp, vp, C.POINTER(C.c_char_p), C.c_uint32]
    L.usb_udb_probe.argtypes = [C.c_char_p]
    L.usb_udb_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.usb_udb_free.argtypes = [vp]
    L.usb_udb_free.restype = None
    L.usb_udb_seq_count.argtypes = [vp]
    L.usb_udb_seq_count.restype = C.c_uint32
    L.usb_udb_is_nucleo.argtypes = [vp]
    L.usb_udb_word_length.argtypes = [vp]
    L.usb_udb_word_length.restype = C.c_uint32
    L.usb_udb_seqs.argtypes = [vp, C.POINTER(u64p)]
    L.usb_udb_seqs.restype = u8p
    L.usb_udb_label.argtypes = [vp, C.c_uint32]
    L.usb_udb_label.restype = C.c_char_p
    L.usb_udb_row.argtypes = [vp, C.c_uint32, C.POINTER(u32p), u32p]
    L.usb_debug_half_row.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, u32p, u32p]
    L.usb_index_row.argtypes = [vp, C.c_uint32, C.POINTER(u32p), u32p]
    L.usb_index_seq.argtypes = [vp, C.c_uint32, C.POINTER(u8p), u32p]
    L.usb_searcher_create.argtypes = [vp, C.POINTER(Params), C.POINTER(vp)]
    L.usb_searcher_free.argtypes = [vp]
    L.usb_searcher_free.restype = None
    L.usb_search_batch.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(vp)]
    L.usb_batch_upload.argtypes = [vp, vp, vp, C.c_uint32]
    L.usb_batch_run.argtypes = [vp, C.POINTER(C.c_float)]
    L.usb_batch_download.argtypes = [vp, C.POINTER(vp)]
    L.usb_batch_counters.argtypes = [vp, u64p]
    L.usb_batch_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.usb_cluster_round.argtypes = [vp, vp, vp, C.c_uint32, u32p, vp, C.POINTER(vp)]
    L.usb_searcher_launch_count.argtypes = [vp]
    L.usb_searcher_launch_count.restype = C.c_uint64
    L.usb_batch_export_hits_device.argtypes = [vp, vp, C.c_uint64, u64p]
    L.usb_result_hit_count.argtypes = [vp]
    L.usb_result_hit_count.restype = C.c_uint64
    L.usb_result_hits.argtypes = [vp]
    L.usb_result_hits.restype = vp
    L.usb_result_runs.argtypes = [vp, u64p]
    L.usb_result_runs.restype = vp
    L.usb_result_query_offsets.argtypes = [vp]
    L.usb_result_query_offsets.restype = vp
    L.usb_result_qstats.argtypes = [vp]
    L.usb_result_qstats.restype = vp
    L.usb_result_free.argtypes = [vp]
    L.usb_result_free.restype = None
    L.usb_result_path.argtypes = [vp, vp, C.c_char_p]
    L.usb_result_path.restype = C.c_uint32
    L.usb_rank_batch.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    L.usb_align_pairs.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, C.POINTER(vp), vp, C.c_uint32]
    L.usb_viterbi_batch.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint32, vp, vp, vp]
    L.usb_set_local.argtypes = [C.POINTER(Params), C.c_int, C.c_float]
    L.usb_set_local.restype = None
    L.usb_index_set_attrs.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp]
    L.usb_batch_set_query_attrs.argtypes = [vp, C.c_uint32, vp, vp]
    L.usb_set_amino.argtypes = [C.POINTER(Params)]
    L.usb_set_amino.restype = None
    L.usb_local_evalue.argtypes = [vp, C.c_int32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.usb_params_evalue.argtypes = [C.POINTER(Params), C.c_int32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.usb_local_pairs.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32, C.POINTER(vp)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise UsbError("usb200 error %d: %s" % (rc, lib().usb_last_error().decode()))


def default_params(cluster_fast=False, **kw):
    """usb_default_params + overrides.  Accepter / Terminator options (maxid, mincols, query_cov, termid ...)
    also set their accept_flags bit; self / notself / selfid are flags only (pass True)."""
    p = Params()
    lib().usb_default_params(C.byref(p), int(cluster_fast))
    for k, v in kw.items():
        if k in ("self", "notself", "selfid"):
            if v:
                p.accept_flags |= ACC[k]
            continue
        if not hasattr(p, k):
            raise AttributeError("unknown usb_params field %r" % k)
        setattr(p, k, v)
        if k in ACC:
            p.accept_flags |= ACC[k]
    return p


def pack_seqs(seqs):
    """list of bytes/str -> (uint8 concatenation, uint64 offsets[n+1])."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if data.size == 0:
        data = np.zeros(1, np.uint8)
    return data, off


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Result:
    """A usb_result (hits grouped by query in the reference's output order).  copy=True (default):
    numpy copies, the handle is released at once.  copy=False: the arrays are views of the library's
    buffers, valid while this object is alive (what a C caller of usb_search_batch sees)."""

    def __init__(self, handle, n_groups, n_jobs, copy=True):
        L = lib()
        self._handle = handle
        n = L.usb_result_hit_count(handle)

        def view(ptr, count, dtype):
            if not count or not ptr:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(int(ptr))
            a = np.frombuffer(buf, dtype=dtype, count=count)
            return a.copy() if copy else a

        self.hits = view(L.usb_result_hits(handle), n, HIT_DTYPE)
        nr = C.c_uint64()
        rp = L.usb_result_runs(handle, C.byref(nr))
        self.runs = view(rp, nr.value, np.uint32)
        self.qoff = view(L.usb_result_query_offsets(handle), n_groups + 1, np.uint64)
        self.qstat = view(L.usb_result_qstats(handle), n_jobs, QSTAT_DTYPE)
        if copy:
            self.close()

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            lib().usb_result_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def path(self, hit):
        ops = "MDI?"
        r = self.runs[int(hit["run_off"]):int(hit["run_off"]) + int(hit["run_cnt"])]
        return "".join(ops[int(v) & 3] * (int(v) >> 2) for v in r)

    def cigar(self, hit):
        """Compressed path as the reference prints it (comppath.cpp:7-48): count omitted when 1."""
        ops = "MDI?"
        r = self.runs[int(hit["run_off"]):int(hit["run_off"]) + int(hit["run_cnt"])]
        return "".join(("%d" % (int(v) >> 2) if (int(v) >> 2) != 1 else "") + ops[int(v) & 3] for v in r)


def udb_write(path, labels, seqs, params=None):
    """-makeudb_usearch (makeudb.cpp:27-60): masks, indexes and writes a .udb file; host only."""
    p = params or default_params()
    data, off = pack_seqs(seqs)
    arr = (C.c_char_p * len(labels))(*[l.encode() if isinstance(l, str) else l for l in labels])
    check(lib().usb_udb_write(os.fsencode(path), C.byref(p), _ptr(data), _ptr(off), arr, len(seqs)))


class Udb:
    """A .udb file in memory (UDBData::FromUDBFile, udbio.cpp:242-279); host only."""

    def __init__(self, path):
        h = C.c_void_p()
        check(lib().usb_udb_read(os.fsencode(path), C.byref(h)))
        self.handle = h
        self.n_seq = lib().usb_udb_seq_count(h)
        self.is_nucleo = bool(lib().usb_udb_is_nucleo(h))
        self.word_length = lib().usb_udb_word_length(h)
        offp = C.POINTER(C.c_uint64)()
        sp = lib().usb_udb_seqs(h, C.byref(offp))
        off = np.ctypeslib.as_array(offp, shape=(self.n_seq + 1,)).copy()
        buf = bytes(np.ctypeslib.as_array(sp, shape=(int(off[-1]),))) if off[-1] else b""
        self.seqs = [buf[int(off[i]):int(off[i + 1])] for i in range(self.n_seq)]
        self.labels = [lib().usb_udb_label(h, i).decode() for i in range(self.n_seq)]

    def row(self, word):
        p = C.POINTER(C.c_uint32)()
        n = C.c_uint32()
        check(lib().usb_udb_row(self.handle, word, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def close(self):
        if self.handle:
            lib().usb_udb_free(self.handle)
            self.handle = None

    def __del__(self):
        self.close()


class Index:
    def __init__(self, seqs, params=None, device=0):
        self.params = params or default_params()
        self._data, self._off = pack_seqs(seqs)
        h = C.c_void_p()
        check(lib().usb_index_create(device, C.byref(self.params), _ptr(self._data), _ptr(self._off), len(seqs),
                                     C.byref(h)))
        self.handle = h
        self.n_seq = len(seqs)

    def append(self, seqs):
        """UDBData::AddSIToDB_CopyData for a block of new targets (cluster_fast centroids)."""
        data, off = pack_seqs(seqs)
        check(lib().usb_index_append(self.handle, _ptr(data), _ptr(off), len(seqs)))
        self.n_seq += len(seqs)

    def row(self, word):
        p = C.POINTER(C.c_uint32)()
        n = C.c_uint32()
        check(lib().usb_index_row(self.handle, word, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def seq(self, t):
        p = C.POINTER(C.c_uint8)()
        n = C.c_uint32()
        check(lib().usb_index_seq(self.handle, t, C.byref(p), C.byref(n)))
        return bytes(np.ctypeslib.as_array(p, shape=(n.value,))) if n.value else b""

    @property
    def posting_count(self):
        return lib().usb_index_posting_count(self.handle)

    @property
    def posting_width(self):
        """Bytes per posting as laid out in HBM now (2 = bank-aware 2-byte rows, 4 = ascending rows)."""
        return lib().usb_index_posting_width(self.handle)

    def close(self):
        if self.handle:
            lib().usb_index_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Searcher:
    def __init__(self, index, params=None):
        self.index = index
        self.params = params or index.params
        h = C.c_void_p()
        check(lib().usb_searcher_create(index.handle, C.byref(self.params), C.byref(h)))
        self.handle = h
        self.strands = 2 if self.params.strand_both else 1

    def search(self, seqs):
        data, off = pack_seqs(seqs)
        return self.search_packed(data, off)

    def search_packed(self, data, off, copy=True):
        n = len(off) - 1
        h = C.c_void_p()
        check(lib().usb_search_batch(self.handle, _ptr(data), _ptr(off), n, C.byref(h)))
        return Result(h, n, n * self.strands, copy)

    def upload(self, data, off):
        self._nq = len(off) - 1
        check(lib().usb_batch_upload(self.handle, _ptr(data), _ptr(off), self._nq))

    def run(self):
        ms = (C.c_float * 3)()
        check(lib().usb_batch_run(self.handle, ms))
        return [ms[0], ms[1], ms[2]]

    def download(self):
        h = C.c_void_p()
        check(lib().usb_batch_download(self.handle, C.byref(h)))
        return Result(h, self._nq, self._nq * self.strands)

    def export_hits_device(self, dev_ptr, cap_hits):
        n = C.c_uint64()
        check(lib().usb_batch_export_hits_device(self.handle, C.c_void_p(dev_ptr), cap_hits, C.byref(n)))
        return n.value

    def cluster_round(self, data, off):
        """One round of the cluster_fast loop on packed queries -> (n_committed, cluster_idx, Result)."""
        n = len(off) - 1
        ncom = C.c_uint32()
        cidx = np.zeros(max(n, 1), np.uint32)
        h = C.c_void_p()
        check(lib().usb_cluster_round(self.handle, _ptr(data), _ptr(off), n, C.byref(ncom), _ptr(cidx), C.byref(h)))
        self.index.n_seq = lib().usb_index_seq_count(self.index.handle)
        return ncom.value, cidx[:ncom.value], Result(h, ncom.value, ncom.value)

    def counters(self):
        out = (C.c_uint64 * 4)()
        check(lib().usb_batch_counters(self.handle, out))
        return dict(postings=out[0], hits=out[1], runs=out[2], jobs=out[3])

    def kernel_ms(self):
        out = (C.c_double * 8)()
        check(lib().usb_batch_kernel_ms(self.handle, out))
        return dict(rank=out[0], gate=out[1], dp=out[2], misc=out[3], dp_records=int(out[4]), dp_cells=int(out[5]),
                    dp_seq_bytes=int(out[6]), hsp_words=int(out[7]))

    @property
    def launch_count(self):
        return lib().usb_searcher_launch_count(self.handle)

    def rank(self, seqs, k_max, want_u=False):
        data, off = pack_seqs(seqs)
        nj = len(seqs) * self.strands
        ct = np.zeros((nj, k_max), np.uint32)
        cu = np.zeros((nj, k_max), np.uint32)
        nc = np.zeros(nj, np.uint32)
        u = np.zeros((nj, self.index.n_seq), np.uint32) if want_u else None
        check(lib().usb_rank_batch(self.handle, _ptr(data), _ptr(off), len(seqs), k_max, _ptr(ct), _ptr(cu), _ptr(nc),
                                   _ptr(u)))
        return ct, cu, nc, u

    def local_evalue(self, raw, ql):
        """(E-value, bit score) of a local hit's raw score (estats.cpp:73-96)."""
        ev, bits = C.c_double(), C.c_double()
        check(lib().usb_local_evalue(self.handle, int(raw), int(ql), C.byref(ev), C.byref(bits)))
        return ev.value, bits.value

    def local_pairs(self, seqs, pair_q, pair_t):
        """LocalAligner2::AlignMulti on explicit (query, target) pairs -> Result grouped by pair."""
        data, off = pack_seqs(seqs)
        pq = np.ascontiguousarray(pair_q, dtype=np.uint32)
        pt = np.ascontiguousarray(pair_t, dtype=np.uint32)
        h = C.c_void_p()
        check(lib().usb_local_pairs(self.handle, _ptr(data), _ptr(off), len(seqs), _ptr(pq), _ptr(pt), len(pq), C.byref(h)))
        return Result(h, len(pq), 0)

    def align_pairs(self, seqs, pair_q, pair_t, max_hsp=0):
        data, off = pack_seqs(seqs)
        pq = np.ascontiguousarray(pair_q, dtype=np.uint32)
        pt = np.ascontiguousarray(pair_t, dtype=np.uint32)
        n = len(pq)
        aligned = np.zeros(max(n, 1), np.uint8)
        hsp = np.zeros((max(n, 1), 1 + 4 * max_hsp), np.uint32) if max_hsp else None
        h = C.c_void_p()
        check(lib().usb_align_pairs(self.handle, _ptr(data), _ptr(off), len(seqs), _ptr(pq), _ptr(pt), n,
                                    _ptr(aligned), C.byref(h), _ptr(hsp), max_hsp))
        return aligned[:n], Result(h, n, 0), hsp

    def viterbi(self, a_seqs, b_seqs, flags):
        a, ao = pack_seqs(a_seqs)
        b, bo = pack_seqs(b_seqs)
        n = len(a_seqs)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        lens = (ao[1:] - ao[:-1]) + (bo[1:] - bo[:-1]) + 1
        po = np.zeros(n, np.uint64)
        if n > 1:
            po[1:] = np.cumsum(lens[:-1])
        paths = np.zeros(int(lens.sum()) + 1, np.uint8)
        sc = np.zeros(max(n, 1), np.int32)
        check(lib().usb_viterbi_batch(self.handle, _ptr(a), _ptr(ao), _ptr(b), _ptr(bo), _ptr(fl), n, _ptr(paths),
                                      _ptr(po), _ptr(sc)))
        out = []
        raw = paths.tobytes()
        for i in range(n):
            s = raw[int(po[i]):int(po[i]) + int(lens[i])]
            out.append(s[:s.index(b"\0")].decode())
        return out, sc[:n]

    def close(self):
        if self.handle:
            lib().usb_searcher_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
