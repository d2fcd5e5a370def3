#!/usr/bin/env python3
"""bench.py -- usearch_global hot-path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--reads R] [--db D] [--no-cpu-baseline] [--no-legs]

Workload (BASELINE.json metric, config 4): usearch_global of R = 1M synthetic 250 bp reads against a
D = 100k x 1500 bp synthetic 16S-like DB at -id 0.97 -strand plus.  One "step" = one pass of the hot
path (U-sort kernel, HSP-gate kernels, DP kernels, commit) over the reads of this rank.

N > 1 (torchrun, one rank per GPU): STRONG scaling -- the same R reads are cut into N contiguous
shards (usearch12_b200.shard.shard_range), every rank holds a replica of the index, and after each
step the hit records that exist (n_hits x 80 bytes, not the buffer capacity) are gathered on rank 0
with NCCL.  A weak-scaling pass (every rank searches all R reads) is reported under "weak".

value  = reads/s with the reads already resident in HBM (CUDA-event time of the kernels on the
         library stream plus the NCCL gather, max over ranks).
e2e    = reads/s through usb_search_batch with pinned HOST buffers: H2D of the reads, kernels,
         D2H of hits/paths/counters, host-side grouping into HitMgr order and, for N > 1, the gather.
legs   = (N = 1) the other BASELINE configs through the same library: config 2 (100k reads vs 10k
         DB), config 3 (cluster_fast, 1M window-random and 1M amplicon reads, host CLI), config 5
         (usearch_local, 200k x 400 aa vs 50k proteins) and the whole -usearch_global command of the
         host CLI on config 4 (FASTA bytes in -> .uc + .b6 bytes out), each next to the reference
         binary on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "query-seqs/sec usearch_global 1Mx250bp vs 100k-seq DB @97%id"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "usearch12")
ORACLE_CLI = os.path.join(ROOT, "oracle", "_build", "uso_cli")


def workload_name(a):
    return "usearch_global %dx250bp synthetic reads vs %dx1500bp synthetic DB, -id 0.97 -strand plus" % (a.reads, a.db)


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU while the timed region runs: NVML in a thread of
    this process (an nvidia-smi child per rank takes a second to start and holds driver locks that
    stretch an 80 ms step at 8 GPUs by ~15 ms); nvidia-smi -lms only where pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    _nvml = None

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []      # nvidia-smi lines
        self.samples = []   # (sm MHz, max MHz, reasons bitmask) from NVML
        self.proc = None
        self.handle = None
        self.stop_flag = threading.Event()
        try:
            import pynvml
            if ClockSampler._nvml is None:
                pynvml.nvmlInit()
                ClockSampler._nvml = pynvml
            nv = ClockSampler._nvml
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:
                self.handle = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def _sample(self):
        nv = ClockSampler._nvml
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
            try:
                why = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                why = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((sm, self.max_mhz, why))
        except Exception:
            pass

    def _loop(self):
        while not self.stop_flag.is_set():
            self._sample()
            self.stop_flag.wait(0.02)

    def start(self):
        if self.handle is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.handle is not None:
            self._sample()  # at least one sample inside the region, however short it was
            self.stop_flag.set()
            self.thread.join(timeout=2)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            sm = [x[0] for x in self.samples]
            reasons = sorted(n for n in names if any(x[2] & bits[n] for x in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz if sm else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------- reference arm
def _run(cmd, **kw):
    t = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, **kw)
    return time.perf_counter() - t


class ReferenceRunner:
    """Times the reference's own CPU implementation (oracle/_ref/usearch12, the unmodified binary
    built by oracle/Makefile.ref from /root/reference/src, gcc -O3 -ffast-math -march=x86-64-v3) --
    or, where that binary is absent, the oracle port -- on a bounded sample of the same workload
    with every host core.  Search time excludes the DB load like the reference's own "Search time"
    log line (search.cpp:128-134): a .udb is prebuilt once and the wall time of a 1-query run
    (load only) is subtracted."""

    def __init__(self, db, db_off, reads, r_off, n_sample, first=0):
        import synth_np
        self.tmp = tempfile.mkdtemp(prefix="usb_ref_")
        self.n = n_sample
        self.dbfa = os.path.join(self.tmp, "db.fa")
        self.qfa = os.path.join(self.tmp, "q.fa")
        self.q1 = os.path.join(self.tmp, "q1.fa")
        synth_np.write_fasta(self.dbfa, db, db_off, "db")
        synth_np.write_fasta(self.qfa, reads, r_off, "q", first, first + n_sample)
        synth_np.write_fasta(self.q1, reads, r_off, "q", first, first + 1)
        self.cores = os.cpu_count() or 1
        if os.path.exists(REF_BIN):
            self.kind = "reference"
            self.udb = os.path.join(self.tmp, "db.udb")
            _run([REF_BIN, "-makeudb_usearch", self.dbfa, "-output", self.udb, "-quiet"])
            self.t_load = min(self._ref(self.q1) for _ in range(2))
        else:
            self.kind = "port"
            self.cores = 1
            self.t_load = self._port(self.q1)

    def _ref(self, q):
        return _run([REF_BIN, "-usearch_global", q, "-db", self.udb, "-id", "0.97", "-strand", "plus", "-threads",
                     str(self.cores), "-uc", os.path.join(self.tmp, "o.uc"), "-quiet"])

    def _port(self, q):
        o = os.path.join(self.tmp, "o")
        return _run([ORACLE_CLI, "usearch_global", q, self.dbfa, "0.97", "plus", o + ".user", o + ".uc", o + ".b6"])

    def step(self):
        """-> reads/s of one sample pass."""
        t = self._ref(self.qfa) if self.kind == "reference" else self._port(self.qfa)
        return self.n / max(t - self.t_load, 1e-3)

    def describe(self):
        return "%d reads of the workload vs the full DB, %s, %d threads, DB load (%.1fs) subtracted" % (
            self.n, "oracle/_ref/usearch12 -usearch_global (built -march=x86-64-v3)" if self.kind == "reference"
            else "oracle/uso_cli (1 thread)", self.cores, self.t_load)

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------- helpers
def make_index(capi, p, db, db_off, n, device):
    ix = capi.Index.__new__(capi.Index)
    ix.params, ix._data, ix._off, ix.n_seq = p, db, db_off, n
    h = C.c_void_p()
    capi.check(capi.lib().usb_index_create(device, C.byref(p), db.ctypes.data_as(C.c_void_p),
                                           db_off.ctypes.data_as(C.c_void_p), n, C.byref(h)))
    ix.handle = h
    return ix


def s_kmax(p, n_db):
    if p.maxaccepts > 0 and p.maxrejects > 0:
        return min(n_db, p.maxaccepts + p.maxrejects - 1)
    return n_db


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def algorithmic_bytes(p, n_db, reads_bytes, pw, ctr, qs, hits, n_runs, km):
    """Per-launch algorithmic bytes of the three kernel families."""
    b_rank = pw * ctr["postings"] + 8.0 * float(np.minimum(qs["n_cand"], s_kmax(p, n_db)).sum()) + float(reads_bytes)
    # gate: packed target letters of every attempt (2 bits each), the query once per stage it takes
    # part in, one verdict word per attempt, a record + chained HSP coordinates per survivor
    tried = qs["n_tried"].astype(np.float64)
    stages = (tried > 0).astype(np.float64) + (tried > 1).astype(np.float64)
    b_gate = float(qs["seq_bytes"].sum()) / 4.0 + float(reads_bytes) * float(stages.mean()) + 4.0 * float(tried.sum()) + \
        32.0 * km["dp_records"] + 4.0 * km["hsp_words"]
    # dp: letters of the record, one trace nibble per cell, trace read + path out along the path,
    # the hit record and its runs
    b_dp = float(km["dp_seq_bytes"]) + float(np.ceil(km["dp_cells"] / 2.0)) + 2.0 * float(km["dp_seq_bytes"]) + \
        80.0 * len(hits) + 4.0 * n_runs
    return {"k_rank": b_rank, "k_gate": b_gate, "k_dp": b_dp}


# ---------------------------------------------------------------------------------- legs (N = 1)
def leg_config2(capi, a):
    """usearch_global 100k x 250 bp reads vs 10k x 1500 bp DB (BASELINE config 2)."""
    import synth_np
    db, db_off = synth_np.gen_db(10000, 1500, seed=1)
    reads, r_off, _ = synth_np.gen_reads(db, db_off, 100000, 250, seed=1001)
    p = capi.default_params()
    s = capi.Searcher(make_index(capi, p, db, db_off, 10000, 0), p)
    s.upload(reads, r_off)
    for _ in range(3):
        s.run()
    ms = [s.run() for _ in range(3)]
    for _ in range(2):  # warm-up of the end-to-end path (staging buffers, recycled result objects)
        res = s.search_packed(reads, r_off, copy=False)
    t = time.perf_counter()
    for _ in range(3):
        res = s.search_packed(reads, r_off, copy=False)
    e2e = (time.perf_counter() - t) / 3
    out = {"workload": "usearch_global 100000x250bp vs 10000x1500bp DB, -id 0.97 -strand plus",
           "value": 100000 / (np.mean([m[2] for m in ms]) / 1e3), "unit": "query-seqs/s",
           "e2e": 100000 / e2e, "hits": int(len(res.hits))}
    if not a.no_cpu_baseline:
        rr = ReferenceRunner(db, db_off, reads, r_off, 50000)
        out["cpu_baseline"] = {"value": rr.step(), "unit": "query-seqs/s", "cores": rr.cores, "kind": rr.kind,
                               "sample": rr.describe()}
        rr.close()
    return out


def leg_cluster(a, amplicon, db, db_off):
    """cluster_fast on 1M x 250 bp reads (BASELINE config 3) through the host CLI, whole command."""
    import synth_np
    from usearch12_b200 import build
    cli = build.build_cli()
    n = a.cluster_reads
    reads, r_off, _ = synth_np.gen_reads(db, db_off, n, 250, seed=3000, window=(500, 750) if amplicon else None)
    tmp = tempfile.mkdtemp(prefix="usb_cl_")
    try:
        fa = os.path.join(tmp, "r.fa")
        synth_np.write_fasta(fa, reads, r_off, "r")
        t = time.perf_counter()
        r = subprocess.run([cli, "-cluster_fast", fa, "-id", "0.97", "-uc", os.path.join(tmp, "o.uc"), "-centroids",
                            os.path.join(tmp, "o.fa")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           env=dict(os.environ, USB_TIMING="1"))
        ours = time.perf_counter() - t
        out = {"workload": "cluster_fast %dx250bp %s reads -id 0.97 (host CLI, FASTA in -> .uc + centroids out)" % (
            n, "amplicon" if amplicon else "window-random"), "value": n / ours, "unit": "seqs/s", "seconds": ours,
            "rc": r.returncode, "log": r.stdout.strip().splitlines()[-3:]}
        if not a.no_cpu_baseline and os.path.exists(REF_BIN):
            ns = min(a.cluster_ref_sample, n)
            fs = os.path.join(tmp, "s.fa")
            synth_np.write_fasta(fs, reads, r_off, "r", 0, ns)
            tr = _run([REF_BIN, "-cluster_fast", fs, "-id", "0.97", "-threads", "1", "-uc", os.path.join(tmp, "r.uc"), "-quiet"])
            subprocess.run([cli, "-cluster_fast", fs, "-id", "0.97", "-uc", os.path.join(tmp, "s.uc"), "-quiet"], check=True)
            same = open(os.path.join(tmp, "s.uc")).read() == open(os.path.join(tmp, "r.uc")).read()
            out["cpu_baseline"] = {"value": ns / tr, "unit": "seqs/s", "cores": 1, "kind": "reference",
                                   "sample": "first %d reads, oracle/_ref/usearch12 -cluster_fast -threads 1 (its search loop is "
                                             "serial, clusterfast.cpp:120-129), whole command" % ns,
                                   "sample_uc_identical_to_ours": same}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def leg_local(capi, a):
    """usearch_local 200k x 400 aa vs 50k proteins, -id 0.5 -evalue 1e-5 (BASELINE config 5)."""
    import synth_np
    nq, ndb = a.local_queries, 50000
    db, doff, q, qoff, _ = synth_np.gen_aa(ndb, 400, nq, 7)
    p = capi.default_params(id=0.5)
    capi.lib().usb_set_local(C.byref(p), 0, 1e-5)
    s = capi.Searcher(make_index(capi, p, db, doff, ndb, 0), p)
    s.upload(q, qoff)
    for _ in range(2):
        s.run()
    ms = [s.run() for _ in range(3)]
    for _ in range(2):  # warm-up of the end-to-end path (staging buffers, recycled result objects)
        res = s.search_packed(q, qoff, copy=False)
    t = time.perf_counter()
    for _ in range(3):
        res = s.search_packed(q, qoff, copy=False)
    e2e = (time.perf_counter() - t) / 3
    out = {"workload": "usearch_local %dx400aa vs %d-seq DB, -id 0.5 -evalue 1e-5" % (nq, ndb),
           "value": nq / (np.mean([m[2] for m in ms]) / 1e3), "unit": "query-seqs/s", "e2e": nq / e2e,
           "kernels_ms": {"k_rank": float(np.mean([m[0] for m in ms])), "k_local": float(np.mean([m[1] for m in ms]))},
           "hits": int(len(res.hits))}
    if not a.no_cpu_baseline and os.path.exists(REF_BIN):
        tmp = tempfile.mkdtemp(prefix="usb_loc_")
        try:
            n = min(20000, nq)
            synth_np.write_fasta(os.path.join(tmp, "db.fa"), db, doff, "p")
            synth_np.write_fasta(os.path.join(tmp, "q.fa"), q, qoff, "q", 0, n)
            synth_np.write_fasta(os.path.join(tmp, "q1.fa"), q, qoff, "q", 0, 1)
            cores = os.cpu_count() or 1
            udb = os.path.join(tmp, "db.udb")
            _run([REF_BIN, "-makeudb_usearch", os.path.join(tmp, "db.fa"), "-output", udb, "-quiet"])

            def ref(qf):
                return _run([REF_BIN, "-usearch_local", qf, "-db", udb, "-id", "0.5", "-evalue", "1e-5", "-threads", str(cores),
                             "-blast6out", os.path.join(tmp, "b6"), "-quiet"])
            load = min(ref(os.path.join(tmp, "q1.fa")) for _ in range(2))
            wall = ref(os.path.join(tmp, "q.fa"))
            out["cpu_baseline"] = {"value": n / max(wall - load, 1e-3), "unit": "query-seqs/s", "cores": cores, "kind": "reference",
                                   "sample": "%d queries vs the full DB, oracle/_ref/usearch12 -usearch_local, %d threads, DB "
                                             "load (%.2fs, 1-query run) subtracted from %.2fs" % (n, cores, load, wall)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return out


def leg_cli(a, db, db_off, reads, r_off):
    """Whole -usearch_global command of the host CLI on config 4: FASTA bytes in -> .uc + .b6 bytes
    out.  The CLI reports its own phases with USB_TIMING=1."""
    import synth_np
    from usearch12_b200 import build
    cli = build.build_cli()
    tmp = tempfile.mkdtemp(prefix="usb_cli_")
    try:
        dbfa, qfa = os.path.join(tmp, "db.fa"), os.path.join(tmp, "q.fa")
        synth_np.write_fasta(dbfa, db, db_off, "db")
        synth_np.write_fasta(qfa, reads, r_off, "q")
        best = None
        for _ in range(2):
            t = time.perf_counter()
            r = subprocess.run([cli, "-usearch_global", qfa, "-db", dbfa, "-id", "0.97", "-strand", "plus", "-uc",
                                os.path.join(tmp, "o.uc"), "-blast6out", os.path.join(tmp, "o.b6")],
                               env=dict(os.environ, USB_TIMING="1"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            wall = time.perf_counter() - t
            m = re.search(r"search\+output ([\d.]+)s", r.stdout)
            rec = {"whole_command_s": wall, "search_output_s": float(m.group(1)) if m else None, "rc": r.returncode,
                   "log": [x for x in r.stdout.splitlines() if x.startswith("timing:")][-1:]}
            if best is None or wall < best["whole_command_s"]:
                best = rec
        n = len(r_off) - 1
        out = {"workload": workload_name(a) + " (host CLI: FASTA in -> .uc + .b6 out)",
               "value": n / best["search_output_s"] if best["search_output_s"] else None, "unit": "query-seqs/s",
               "what": "queries / (search + output phase: first batch H2D -> last output byte written), index build excluded "
                       "like the reference's Search time", "whole_command_reads_per_s": n / best["whole_command_s"]}
        out.update(best)
        out["out_bytes"] = os.path.getsize(os.path.join(tmp, "o.uc")) + os.path.getsize(os.path.join(tmp, "o.b6"))
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=1000000)
    ap.add_argument("--db", type=int, default=100000)
    ap.add_argument("--ref-sample", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true")
    ap.add_argument("--legs", default="config2,cluster,cluster_amplicon,local,cli")
    ap.add_argument("--cluster-reads", type=int, default=1000000)
    ap.add_argument("--cluster-ref-sample", type=int, default=120000)  # past the -big switch (100 000 clusters)
    ap.add_argument("--local-queries", type=int, default=200000)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else max(a.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import synth_np
    if a.impl == "reference":
        if rank != 0:
            return 0
        n = min(a.ref_sample, a.reads)
        db, db_off = synth_np.gen_db(a.db, 1500, seed=4)
        reads, r_off, _ = synth_np.gen_reads(db, db_off, n, 250, seed=1000)
        rr = ReferenceRunner(db, db_off, reads, r_off, n)
        for _ in range(a.warmup):
            rr.step()
        t0 = time.perf_counter()
        vals = [rr.step() for _ in range(a.steps)]
        wall = time.perf_counter() - t0
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "query-seqs/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * wall / max(1, a.steps),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic", "config": {"workload": workload_name(a)},
                "cpu_baseline": {"value": v, "unit": "query-seqs/s", "cores": rr.cores, "kind": rr.kind,
                                 "sample": rr.describe()},
                "e2e": {"value": v, "unit": "query-seqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        rr.close()
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from usearch12_b200 import capi
    from usearch12_b200.shard import shard_range
    if not torch.cuda.is_available() or capi.lib().usb_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the usb200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic workload: replicated DB; the SAME reads on every rank, sharded by rank
    t_gen = time.perf_counter()
    db, db_off = synth_np.gen_db(a.db, 1500, seed=4)
    reads, r_off, _ = synth_np.gen_reads(db, db_off, a.reads, 250, seed=1000)
    t_gen = time.perf_counter() - t_gen
    lo, hi = shard_range(a.reads, rank, world)
    n_local = hi - lo
    p = capi.default_params()
    t_ix = time.perf_counter()
    ix = make_index(capi, p, db, db_off, a.db, local_rank)
    t_ix = time.perf_counter() - t_ix
    s = capi.Searcher(ix, p)

    # pinned host copies of the inputs (the whole read set: the weak pass uses all of it)
    pin_reads = torch.empty(reads.size, dtype=torch.uint8, pin_memory=True)
    pin_reads.numpy()[:] = reads
    pin_off = torch.empty(r_off.size, dtype=torch.int64, pin_memory=True)
    pin_off.numpy()[:] = r_off.astype(np.int64)
    h_reads, h_off_all = pin_reads.numpy(), pin_off.numpy().view(np.uint64)
    h_off = h_off_all[lo:hi + 1]

    HB = capi.HIT_DTYPE.itemsize
    cap_hits = a.reads * max(1, p.maxaccepts)
    gather_src = torch.zeros(cap_hits * HB, dtype=torch.uint8, device="cuda") if world > 1 else None
    gather_dst = torch.zeros(cap_hits * world * HB, dtype=torch.uint8, device="cuda") \
        if world > 1 and rank == 0 else None
    counts_dev = torch.zeros(world, dtype=torch.int64, device="cuda") if world > 1 else None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gather_hits():
        """NCCL gather of the hit records that exist (n_hits x 80 B per rank) onto rank 0; returns
        (device ms, bytes received by rank 0)."""
        n = s.export_hits_device(gather_src.data_ptr(), cap_hits)
        ev0.record()
        mine = torch.tensor([n], dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(counts_dev, mine)
        cnt = counts_dev.cpu().numpy()
        m = int(cnt.max()) * HB
        if rank == 0:
            if m * world > gather_dst.numel():
                raise SystemExit("gather buffer too small")
            outs = [gather_dst[r * m:(r + 1) * m] for r in range(world)]
            dist.gather(gather_src[:m], outs, dst=0)
        else:
            dist.gather(gather_src[:m], None, dst=0)
        ev1.record()
        ev1.synchronize()
        return ev0.elapsed_time(ev1), int(cnt.sum()) * HB

    def device_step():
        ms = s.run()
        g_ms, g_bytes = (0.0, 0)
        if world > 1:
            g_ms, g_bytes = gather_hits()
        return ms, g_ms, g_bytes

    def timed_pass(e2e_off):
        """W warm-up + K timed device steps on the uploaded batch, then K end-to-end steps."""
        sampler = ClockSampler(local_rank)  # (NVML handle set up before the warm-up, outside the timed region)
        for _ in range(a.warmup):
            device_step()
        sync_all()
        sampler.start()
        launches0 = s.launch_count
        t0 = time.perf_counter()
        acc = {"rank": 0.0, "align": 0.0, "gate": 0.0, "dp": 0.0, "misc": 0.0, "tot": 0.0, "gather": 0.0}
        km, g_bytes = None, 0
        for _ in range(a.steps):
            ms, g_ms, g_bytes = device_step()
            km = s.kernel_ms()
            acc["rank"] += ms[0]
            acc["align"] += ms[1]
            acc["tot"] += ms[2] + g_ms
            acc["gather"] += g_ms
            acc["gate"] += km["gate"]
            acc["dp"] += km["dp"]
            acc["misc"] += km["misc"]
        sync_all()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        launches = s.launch_count - launches0
        ctr = s.counters()
        res = s.download()
        # end to end through the C ABI with host buffers (copy=False: the result arrays are the
        # library's own buffers, as a C caller of usb_search_batch gets them)
        s.search_packed(h_reads, e2e_off, copy=False)
        sync_all()
        l1 = s.launch_count
        t1 = time.perf_counter()
        r2 = None
        for _ in range(a.steps):
            r2 = None  # release the previous result first, like a C caller would
            r2 = s.search_packed(h_reads, e2e_off, copy=False)
            if world > 1:
                gather_hits()
        sync_all()
        e2e_wall = time.perf_counter() - t1
        launches += s.launch_count - l1
        h2d = int((int(e2e_off[-1]) - int(e2e_off[0])) + e2e_off.size * 8)
        d2h = int(r2.hits.nbytes + r2.runs.nbytes + r2.qstat.nbytes + 4 * len(r2.qstat) + 24)
        times = torch.tensor([acc["tot"], wall * 1000.0, e2e_wall * 1000.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        tot_ms, wall_ms, e2e_ms = [float(x) for x in times.cpu()]
        return dict(acc=acc, km=km, ctr=ctr, res=res, clocks=clocks, launches=int(launches), tot_ms=tot_ms,
                    wall_ms=wall_ms, e2e_ms=e2e_ms, h2d=h2d, d2h=d2h, g_bytes=g_bytes)

    # ---- strong-scaling pass: this rank's shard of the R reads
    s.upload(h_reads, h_off)
    P1 = timed_pass(h_off)
    value = a.reads * a.steps / (P1["tot_ms"] / 1000.0)
    e2e = a.reads * a.steps / (P1["e2e_ms"] / 1000.0)

    # ---- weak-scaling pass (N > 1): every rank searches all R reads
    weak = None
    if world > 1:
        s.upload(h_reads, h_off_all)
        P2 = timed_pass(h_off_all)
        weak = {"value": a.reads * world * a.steps / (P2["tot_ms"] / 1000.0), "unit": "query-seqs/s",
                "reads_per_gpu": a.reads, "ms_per_step": P2["tot_ms"] / a.steps,
                "e2e": a.reads * world * a.steps / (P2["e2e_ms"] / 1000.0),
                "nccl_gather_ms": P2["acc"]["gather"] / a.steps, "gather_bytes_per_step": P2["g_bytes"]}

    if rank == 0:
        acc, km, res = P1["acc"], P1["km"], P1["res"]
        qs, hits = res.qstat, res.hits
        pw = float(ix.posting_width)  # 2 = bank-aware 2-byte rows (DBs up to 131 070 targets), else 4
        shard_bytes = int(h_off[-1]) - int(h_off[0])
        alg = algorithmic_bytes(p, a.db, shard_bytes, pw, P1["ctr"], qs, hits, len(res.runs), km)
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ms_k = {"k_rank": acc["rank"] / a.steps, "k_gate": acc["gate"] / a.steps, "k_dp": acc["dp"] / a.steps}
        if ms_k["k_gate"] == 0.0:  # one-kernel candidate loop
            ms_k = {"k_rank": acc["rank"] / a.steps, "k_align": acc["align"] / a.steps}
            alg["k_align"] = alg["k_gate"] + alg["k_dp"]
        dom = max(ms_k, key=lambda k: ms_k[k])
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            if prof.get("workload") != workload_name(a) or world != 1:
                prof = {}
        except (OSError, ValueError):
            prof = {}

        def traffic_of(name):
            """DRAM bytes per step from the committed ncu capture -- only if that capture timed the
            kernel within 5 % of what was measured just now (else it describes another binary)."""
            t = prof.get("kernels", {}).get(name)
            if not t or not ms_k.get(name):
                return None
            if abs(t["gpu_time_ms"] - ms_k[name]) > 0.05 * ms_k[name]:
                return None
            return t["dram_bytes"]

        bounds = {
            "k_rank": "hbm during the posting walk, whose rate is set by one shared-memory atomic per posting (see "
                      "shared_atomic_floor); about a third of the launch is the serial scan/select tail",
            "k_gate": "instruction issue (ncu: issue slots 71 % busy in the stage that holds 86 % of its time, DRAM < 1 % of "
                      "peak): integer/bit work on packed letters in shared memory; its HBM fraction is small by construction",
            "k_dp": "shared-memory latency / issue (ncu: issue slots 46-62 % busy at 16 warps/SM); writes one trace byte per "
                    "cell to HBM",
            "k_align": "instruction issue / latency"}

        def shared_atomic_floor():
            """k_rank counts every posting with one shared-memory atomic; tools/ubench_smem.cu measured 3.85 cycles per
            warp-wide random atomic per SM on B200 (profiles/r2_ubench_smem.txt).  Time the launch would take if it did
            nothing but those atomics at that rate on all SMs -- the floor of this counting design."""
            try:
                n_sm = torch.cuda.get_device_properties(0).multi_processor_count
                mhz = float((P1["clocks"] or {}).get("sm_mhz") or 1965.0)
                ms = float(P1["ctr"]["postings"]) * (3.85 / 32.0) / (n_sm * mhz * 1e6) * 1e3
                return {"ms": ms, "frac": ms / ms_k["k_rank"] if ms_k.get("k_rank") else None,
                        "cycles_per_warp_atomic": 3.85, "source": "profiles/r2_ubench_smem.txt"}
            except Exception:  # explanatory key only: never take the line down
                return None

        def kernel_line(name):
            ms = ms_k[name]
            gbps = alg[name] / (ms / 1000.0) / 1e9 if ms else None
            line = {"ms": ms, "share_of_step": ms / max(sum(ms_k.values()), 1e-9), "algorithmic_bytes_per_launch": alg[name],
                    "achieved_GBps": gbps, "frac_of_hbm_peak": gbps / peak if gbps else None, "traffic": traffic_of(name),
                    "bound": bounds[name]}
            if name == "k_rank":
                line["shared_atomic_floor"] = shared_atomic_floor()
            return line
        achieved = alg[dom] / (ms_k[dom] / 1000.0) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "query-seqs/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": P1["tot_ms"] / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(a), "reads_total": a.reads, "reads_per_gpu": n_local, "db_seqs": a.db,
                       "postings": int(ix.posting_count), "posting_bytes": int(pw),
                       "l2": "inputs larger than L2 (reads %d MB + postings %d MB per GPU)" % (
                           shard_bytes >> 20, (int(pw) * ix.posting_count) >> 20),
                       "hit_rate": float(len(hits)) / max(1, n_local), "gen_s": round(t_gen, 1), "index_build_s": round(t_ix, 1)},
            "kernels_ms_per_step": dict(ms_k, k_prep_commit=acc["misc"] / a.steps, nccl_gather=acc["gather"] / a.steps,
                                        dp_records=km["dp_records"], dp_cells=km["dp_cells"], wall=P1["wall_ms"] / a.steps),
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic_of(dom), "peak_source": "measured" if peaks else "fallback",
                         "algorithmic_bytes_per_launch": alg[dom], "kernels": {k: kernel_line(k) for k in ms_k}},
            "e2e": {"value": e2e, "unit": "query-seqs/s", "h2d_bytes_per_step": P1["h2d"], "d2h_bytes_per_step": P1["d2h"],
                    "ms_per_step": P1["e2e_ms"] / a.steps, "gather_bytes_per_step": P1["g_bytes"]},
            "gpu_launches": P1["launches"], "clocks": P1["clocks"],
        }
        if weak:
            line["weak"] = weak
        line["cpu_baseline"] = None
        if world == 1 and not a.no_cpu_baseline:
            n = min(a.ref_sample, a.reads)
            rr = ReferenceRunner(db, db_off, reads, r_off, n)
            v = rr.step()
            line["cpu_baseline"] = {"value": v, "unit": "query-seqs/s", "cores": rr.cores, "kind": rr.kind,
                                    "sample": rr.describe()}
            rr.close()
        if world == 1 and not a.no_legs:
            # the other BASELINE configs; a failing leg is reported, it never takes the main line down
            del s, ix, res, P1
            legs = {}
            for name in [x for x in a.legs.split(",") if x]:
                t = time.perf_counter()
                try:
                    if name == "config2":
                        legs[name] = leg_config2(capi, a)
                    elif name == "cluster":
                        legs[name] = leg_cluster(a, False, db, db_off)
                    elif name == "cluster_amplicon":
                        legs[name] = leg_cluster(a, True, db, db_off)
                    elif name == "local":
                        legs[name] = leg_local(capi, a)
                    elif name == "cli":
                        legs[name] = leg_cli(a, db, db_off, reads, r_off)
                    else:
                        legs[name] = {"error": "unknown leg"}
                except Exception as e:  # noqa: BLE001
                    legs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                legs[name]["leg_seconds"] = round(time.perf_counter() - t, 1)
            line["legs"] = legs
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
