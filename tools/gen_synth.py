#!/usr/bin/env python3
"""Deterministic synthetic nucleotide DB + read generator (spec: SURVEY.md section 8d / Appendix C).

    gen_synth.py NDB DBLEN NQ QLEN SEED PREFIX [NROOT]

Writes PREFIX.db.fa (targets db<i>) and PREFIX.q.fa (reads q<i>;t=db<k>;p=<pos>).
Pure python `random.Random(seed)` so the stream is identical on every box.  This is the
generator the parity fixtures under tests/golden/ were made with; bench.py uses a faster
numpy generator of the same statistical shape for the 1M-read workloads.
"""
import random
import sys


def mutate(s, rate, rng, indel=0.2):
    out = []
    for c in s:
        if rng.random() < rate:
            k = rng.random()
            if k < indel / 2:
                continue  # deletion
            elif k < indel:
                out.append(c)
                out.append(rng.choice("ACGT"))  # insertion after
            else:
                out.append(rng.choice([x for x in "ACGT" if x != c]))  # substitution
        else:
            out.append(c)
    return "".join(out)


def generate(ndb, dblen, nq, qlen, seed, nroot=None):
    rng = random.Random(seed)
    if nroot is None:
        nroot = max(1, ndb // 100)
    roots = ["".join(rng.choice("ACGT") for _ in range(dblen)) for _ in range(nroot)]
    db = [mutate(roots[i % nroot], rng.uniform(0.03, 0.15), rng) for i in range(ndb)]
    reads = []
    for i in range(nq):
        t = rng.randrange(ndb)
        s = db[t]
        p = rng.randrange(0, max(1, len(s) - qlen))
        q = mutate(s[p:p + qlen], rng.uniform(0.0, 0.04), rng)
        if rng.random() < 0.05:
            q = "".join(rng.choice("ACGT") for _ in range(qlen))
        reads.append((">q%d;t=db%d;p=%d" % (i, t, p), q))
    return db, reads


def main(argv):
    ndb, dblen, nq, qlen, seed = map(int, argv[1:6])
    pref = argv[6]
    nroot = int(argv[7]) if len(argv) > 7 else None
    db, reads = generate(ndb, dblen, nq, qlen, seed, nroot)
    with open(pref + ".db.fa", "w") as f:
        f.write("".join(">db%d\n%s\n" % (i, s) for i, s in enumerate(db)))
    with open(pref + ".q.fa", "w") as f:
        for lab, q in reads:
            f.write("%s\n%s\n" % (lab, q))


if __name__ == "__main__":
    main(sys.argv)
