#!/usr/bin/env python3
"""Small fuzz run over the host readers (no GPU): random and malformed FASTA / FASTQ inputs through the ASan + UBSan
build of tools/format_replay.cpp (tools/host_sanitize.py makes it), serial and in pieces; and .udb files with
corrupted headers, bodies and lengths through the plain tool.  A run may end with the reference's error message
(exit code 1) or normally (0); anything else, or a sanitizer message, is a finding.
Usage: python tools/host_sanitize.py && python tools/host_fuzz.py [iterations]"""
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASAN = "/tmp/format_replay_asan"
PLAIN = os.path.join(ROOT, "usearch12_b200", "format_replay")
CLI = os.path.join(ROOT, "usearch12_b200", "usearch12_b200_cli")
D = "/tmp/usb_fuzz"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    os.makedirs(D, exist_ok=True)
    rng = random.Random(7)
    open(D + "/t.fa", "w").write(">t\nACGTACGTAC\n")
    open(D + "/e.tsv", "w").close()
    alphabet = b"ACGTNacgt@>+\n\r -.IIII;=1"
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    findings = 0
    for it in range(n):
        kind = rng.random()
        if kind < 0.4:
            data = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 200)))
        elif kind < 0.7:
            data = b"".join(b"@r%d\n" % i + bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 12))) +
                            rng.choice([b"\n", b"\r\n"]) + b"+\n" + b"I" * rng.randrange(0, 12) + b"\n"
                            for i in range(rng.randrange(1, 5)))
        else:
            data = b"".join(b">r%d\n" % i + bytes(rng.choice(b"ACGT\n-. x") for _ in range(rng.randrange(0, 40))) + b"\n"
                            for i in range(rng.randrange(1, 5)))
        open(D + "/q", "wb").write(data)
        for th in ("1", "3"):
            r = subprocess.run([ASAN, "-query", D + "/q", "-db", D + "/t.fa", "-hits", D + "/e.tsv", "-notmatched", D + "/o"],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(env, USB_FASTA_THREADS=th))
            out = r.stdout.decode(errors="replace")
            if "Sanitizer" in out or "runtime error" in out or r.returncode not in (0, 1):
                findings += 1
                print("FINDING fastx", it, th, r.returncode, out[:400])
    print("fastx: %d inputs x 2, findings so far %d" % (n, findings))
    # .udb files
    open(D + "/small.fa", "w").write("".join(">s%d\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(300))) for i in range(40)))
    subprocess.run([CLI, "-makeudb_usearch", D + "/small.fa", "-output", D + "/small.udb", "-quiet"], check=True)
    base = open(D + "/small.udb", "rb").read()
    open(D + "/q.fa", "w").write(">q\nACGTACGTACGTACGTACGT\n")
    codes = {}
    for it in range(max(100, n * 2 // 3)):
        b = bytearray(base)
        k = rng.random()
        if k < 0.5:
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(0, min(len(b), 400))] = rng.randrange(256)
        elif k < 0.8:
            for _ in range(rng.randrange(1, 20)):
                b[rng.randrange(0, len(b))] = rng.randrange(256)
        else:
            b = b[:rng.randrange(0, len(b))]
        open(D + "/m.udb", "wb").write(bytes(b))
        r = subprocess.run([PLAIN, "-query", D + "/q.fa", "-db", D + "/m.udb", "-hits", D + "/e.tsv"], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT)
        codes[r.returncode] = codes.get(r.returncode, 0) + 1
        if r.returncode not in (0, 1):
            findings += 1
            print("FINDING udb", it, r.returncode, r.stdout.decode(errors="replace")[-300:])
    print("udb: exit codes", codes)
    print("clean" if not findings else "%d findings" % findings)
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
