#!/usr/bin/env python3
"""Config 5 (BASELINE.json configs[4]): usearch_local of synthetic 400-aa proteins vs a 50k-sequence
DB, -id 0.5 -evalue 1e-5, 1xB200, next to the reference binary on the host cores.

    python tools/bench_local.py [--db 50000] [--queries 200000] [--len 400] [--ref-sample 20000]

Prints one JSON line: device-resident and end-to-end query-seqs/s of the CUDA path, the per-kernel
times, and the reference's "Search time" throughput on a sample of the same queries.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", type=int, default=50000)
    ap.add_argument("--queries", type=int, default=200000)
    ap.add_argument("--len", type=int, default=400)
    ap.add_argument("--ref-sample", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=300, help="queries compared hit for hit with the oracle")
    a = ap.parse_args()
    import ctypes as C
    from usearch12_b200 import capi
    db, doff, q, qoff, truth = synth_np.gen_aa(a.db, a.len, a.queries, 7)
    p = capi.default_params(id=0.5)
    capi.lib().usb_set_local(C.byref(p), 0, 1e-5)
    t0 = time.time()
    ix = capi.Index.__new__(capi.Index)
    ix.params = p
    ix._data, ix._off = db, doff
    h = C.c_void_p()
    capi.check(capi.lib().usb_index_create(0, C.byref(p), capi._ptr(db), capi._ptr(doff), a.db, C.byref(h)))
    ix.handle, ix.n_seq = h, a.db
    t_index = time.time() - t0
    s = capi.Searcher(ix, p)
    s.upload(q, qoff)
    ms = None
    for _ in range(1 + a.steps):
        ms = s.run()
    res = s.download()
    t0 = time.time()
    for _ in range(a.steps):
        res = s.search_packed(q, qoff)
    e2e = (time.time() - t0) / a.steps
    nh = len(res.hits)
    top_true = 0
    first = res.qoff[:-1].astype(np.int64)
    has = (res.qoff[1:] > res.qoff[:-1])
    top_true = int((res.hits["target"][first[has]] == truth[has]).sum())
    out = {
        "workload": "usearch_local %dx%daa vs %d-seq DB, -id 0.5 -evalue 1e-5" % (a.queries, a.len, a.db),
        "index_s": round(t_index, 2), "rank_ms": round(ms[0], 2), "local_ms": round(ms[1], 2),
        "device_qps": round(a.queries / (ms[2] / 1e3), 1), "e2e_qps": round(a.queries / e2e, 1),
        "hits": nh, "queries_with_hit": int(has.sum()), "top_hit_is_true_target": top_true,
        "dp_cells": int(res.qstat["dp_cells"].sum()), "gapped_extensions": int(res.qstat["n_dp"].sum()),
        "targets_tried": int(res.qstat["n_tried"].sum()), "postings": s.counters()["postings"],
    }
    if a.check:
        from oracle import uso_py as O
        from tests import util
        n = min(a.check, a.queries)
        dbs = [db[int(doff[i]):int(doff[i + 1])].tobytes() for i in range(a.db)]
        qs = [q[int(qoff[i]):int(qoff[i + 1])].tobytes() for i in range(n)]
        op = util.oracle_local_params(False, id=0.5, evalue=1e-5)
        osr = O.Searcher(O.DB(dbs, op), op)
        lab = ["q%d" % i for i in range(n)]
        dl = ["p%d" % i for i in range(a.db)]
        want = util.oracle_lines_local(osr, lab, qs, dl, False)
        sub = s.search(qs)
        got = util.product_lines_local(sub, s, lab, qs, dl, False)
        out["oracle_check"] = "identical" if got == want else "DIFFERENT"
        out["oracle_checked_queries"] = n
    ref = os.path.join(ROOT, "oracle", "_ref", "usearch12")
    if a.ref_sample and os.path.exists(ref):
        with tempfile.TemporaryDirectory() as td:
            n = min(a.ref_sample, a.queries)
            synth_np.write_fasta(os.path.join(td, "db.fa"), db, doff, "p")
            synth_np.write_fasta(os.path.join(td, "q.fa"), q, qoff, "q", 0, n)
            cores = os.cpu_count()
            log = os.path.join(td, "log")
            t0 = time.time()
            subprocess.run([ref, "-usearch_local", os.path.join(td, "q.fa"), "-db", os.path.join(td, "db.fa"), "-id", "0.5",
                            "-evalue", "1e-5", "-threads", str(cores), "-blast6out", os.path.join(td, "b6"), "-log", log,
                            "-quiet"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            wall = time.time() - t0
            txt = open(log).read()
            m = re.search(r"Search time .*\(([\d.]+)s\)", txt)
            m2 = re.search(r"Db load time .*\(([\d.]+)s\)", txt)
            load = float(m2.group(1)) if m2 else 0.0
            search = max(wall - load, 1e-3)
            out["reference"] = {"cores": cores, "sample_queries": n, "wall_s": round(wall, 2), "db_load_s": load,
                                "search_s_logged": float(m.group(1)) if m else None,
                                "qps": round(n / search, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
