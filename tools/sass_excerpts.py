#!/usr/bin/env python3
"""SASS excerpts of the shipped library for profiles/: the TMA issue/wait sequence of k_gate, the
posting walk of k_rank (2-byte layout) and the four-columns-per-lane row sweep of k_dp.

    tools/sass_excerpts.py [TAG]      -> profiles/TAG_sass_excerpts.txt   (cuobjdump -sass, no GPU needed)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "usearch12_b200", "libusb200.so")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    fn, cur = {}, None
    for l in out:
        if "Function :" in l:
            cur = l.split("Function :")[1].strip()
            fn[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and ";" in l:
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?;)", l)
            if m:
                fn[cur].append("/*%s*/  %s" % (m.group(1), re.sub(r"\s+", " ", m.group(2))))
    return fn


def demangled(name):
    try:
        return subprocess.run(["cu++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip() or name
    except OSError:
        return name


def find(fn, key):
    return [(n, b) for n, b in fn.items() if key in n]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    fn = functions()
    out = ["SASS excerpts of usearch12_b200/libusb200.so (sm_100a), made by tools/sass_excerpts.py from cuobjdump -sass", ""]
    counts = {}
    for n, b in fn.items():
        for op in ("UBLKCP", "SYNCS", "ATOMS", "LDG.E.NA.128", "SHFL.UP", "VIMNMX", "MATCH.ANY", "VOTE", "UTMALDG", "TCGEN05"):
            c = sum(op in l for l in b)
            if c:
                counts.setdefault(demangled(n), {})[op] = c
    out.append("instruction counts per kernel (static):")
    for n, c in counts.items():
        out.append("  %-60s %s" % (n[:60], "  ".join("%s %d" % kv for kv in sorted(c.items()))))
    out.append("")

    for n, b in find(fn, "k_gate"):
        out.append("== %s: bulk copy (TMA, 1-D) of the next candidate's packed letters into shared memory" % demangled(n))
        out.append("   mbarrier.arrive.expect_tx -> SYNCS.ARRIVE.TRANS64, cp.async.bulk -> UBLKCP, mbarrier.try_wait.parity -> SYNCS.PHASECHK")
        ix = [i for i, l in enumerate(b) if "UBLKCP" in l]
        for i in ix[:1]:
            out += ["   " + l for l in b[max(0, i - 16):i + 4]]
        w = [i for i, l in enumerate(b) if "SYNCS.PHASECHK" in l]
        if w:
            out.append("   ...")
            out += ["   " + l for l in b[max(0, w[0] - 3):w[0] + 5]]
        out.append("")

    for n, b in find(fn, "k_rankILb0"):
        ix = [i for i, l in enumerate(b) if "LDG.E.NA.128" in l]
        if not ix:
            continue
        out.append("== %s: posting walk, 2-byte increment descriptors: four 128-bit streaming loads in flight per lane," % demangled(n))
        out.append("   then per 32-bit word two entries: mask/shift to a byte address and one ATOMS each (constant increment per byte class)")
        i0 = ix[0]
        j, atoms = i0, 0
        while j < len(b) and atoms < 16:
            atoms += "ATOMS" in b[j]
            j += 1
        out += ["   " + l for l in b[max(0, i0 - 2):j + 1]]
        out.append("")

    for n, b in find(fn, "k_dp"):
        ld = [i for i, l in enumerate(b) if re.search(r"\bLD[SG]?\S*\.128", l)]
        st = [i for i, l in enumerate(b) if re.search(r"\bST[SG]?\S*\.128", l)]
        pick = None
        for i in ld:
            nxt = [s for s in st if i < s < i + 400]
            if nxt and any("SHFL.UP" in l for l in b[i:nxt[0]]):
                pick = (i, nxt[-1] if nxt[-1] < i + 400 else nxt[0])
                break
        if pick:
            out.append("== %s: one step of the wide row sweep (four band columns per lane): 128-bit loads of the previous" % demangled(n))
            out.append("   row, match state from the packed letters, serial insert state inside the lane, max-plus scan over lanes (SHFL.UP),")
            out.append("   four trace bytes in one 32-bit store, 128-bit stores of the new row")
            out += ["   " + l for l in b[max(0, pick[0] - 4):pick[1] + 3]]
            out.append("")

    for n, b in find(fn, "k_usort_full"):
        ix = [i for i, l in enumerate(b) if "MATCH.ANY" in l]
        if ix:
            out.append("== %s: stable placement step of the whole-list counting sort: MATCH.ANY groups the lanes of one" % demangled(n))
            out.append("   counter value, the group's first lane takes its slots from the shared-memory offset table, SHFL hands the base out")
            i0 = ix[-1]
            out += ["   " + l for l in b[max(0, i0 - 10):i0 + 28]]
            out.append("")

    path = os.path.join(ROOT, "profiles", "%s_sass_excerpts.txt" % tag)
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(path, len(out), "lines")


if __name__ == "__main__":
    main()
