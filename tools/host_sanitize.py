#!/usr/bin/env python3
"""AddressSanitizer + UndefinedBehaviorSanitizer over the host sinks (no GPU): builds tools/format_replay.cpp with
-fsanitize=address,undefined and replays every fixture of tests/test_formats_cpu.py (output formats, FASTQ queries, OTU
table sink, -fastx_uniques writer), serial and threaded formatting.   Usage: python tools/host_sanitize.py
Prints one line per run: name, exit code, number of "runtime error" messages."""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden_formats as M  # noqa: E402
from tests import util  # noqa: E402
from tests.test_formats_cpu import OUT_FLAGS, REPLAY_OPTS, _otutab_inputs, golden_bytes, out_paths  # noqa: E402
from usearch12_b200 import build  # noqa: E402

HOST = os.path.join(ROOT, "usearch12_b200", "csrc", "host")
A = "/tmp/format_replay_asan"


def run(name, cmd, env, cwd=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=cwd)
    n = r.stdout.count("runtime error") + r.stdout.count("AddressSanitizer")
    print("%-12s rc %d, %d sanitizer messages" % (name, r.returncode, n))
    if r.returncode or n:
        print(r.stdout[:3000])
    return r.returncode or n


def main():
    cli = build.build_cli()
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                    "-o", A, os.path.join(ROOT, "tools", "format_replay.cpp"), os.path.join(HOST, "usb_host.cpp"),
                    os.path.join(HOST, "usb_cluster_host.cpp"), "-L" + os.path.join(ROOT, "usearch12_b200"), "-lusb200",
                    "-Wl,-rpath," + os.path.join(ROOT, "usearch12_b200")], check=True)
    bad = 0
    for chunk in ("", "16"):
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
        if chunk:
            env["USB_FORMAT_CHUNK"] = chunk
        for name in M.VARIANTS:
            with tempfile.TemporaryDirectory() as tmp:
                q, d = M.write_inputs(name, tmp)
                udb = os.path.join(tmp, "db.udb")
                subprocess.run([cli, "-makeudb_usearch", d, "-output", udb, "-quiet"], check=True)
                hits = os.path.join(tmp, "h.tsv")
                open(hits, "wb").write(golden_bytes(name, "hits"))
                cmd = [A, "-query", q, "-db", udb, "-hits", hits, "-userfields", M.VARIANTS[name][5]] + REPLAY_OPTS[name]
                for k, p in out_paths(name, tmp).items():
                    cmd += [OUT_FLAGS[k], p]
                bad += bool(run(name + ("/mt" if chunk else ""), cmd, env))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    with tempfile.TemporaryDirectory() as tmp:
        _, d = M.write_inputs("fmt_nt", tmp)
        q = os.path.join(tmp, "q.fq")
        M.write_fastq(q)
        hits = os.path.join(tmp, "h.tsv")
        open(hits, "wb").write(golden_bytes("fmt_nt", "hits"))
        bad += bool(run("fastq", [A, "-query", q, "-db", d, "-hits", hits, "-matchedfq", tmp + "/a", "-notmatchedfq", tmp + "/b"], env))
        _otutab_inputs(tmp)
        open(os.path.join(tmp, "oh.tsv"), "wb").write(golden_bytes("otutab", "hits"))
        bad += bool(run("otutab", [A, "-query", "otutab_reads.fa", "-db", "otutab_otus.fa", "-hits", "oh.tsv", "-otutabout", "t",
                                   "-mapout", "m", "-biomout", "b"], env, cwd=tmp))
        src = os.path.join(tmp, "u.fa")
        open(src, "wb").write(gzip.open(os.path.join(util.GOLDEN, "uniq_in.fa.gz")).read())
        bad += bool(run("uniques", [A, "-uniques", src, "-fastaout", tmp + "/o", "-sizein", "-sizeout", "-topn", "40", "-relabel", "U"], env))
    print("clean" if not bad else "%d runs with findings" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
