#!/usr/bin/env python3
"""Deterministic synthetic protein DB + query generator (spec: SURVEY.md section 8d, config 5).

    gen_synth_aa.py NDB LEN NQ SEED PREFIX [NROOT]

DB: NROOT (default NDB/20) uniform-random 20-letter roots of length LEN; target i = root[i mod NROOT]
mutated per residue with p ~ U(0.05, 0.5).  Queries: a random target mutated with p ~ U(0.05, 0.4).
Mutation mix: 5 % deletion, 5 % insertion of a random residue after, 90 % substitution by a uniformly
random residue.  Writes PREFIX.db.fa (targets p<i>) and PREFIX.q.fa (queries a<i>;t=p<k>).
Pure python `random.Random(seed)`: identical on every box.
"""
import random
import sys

AA = "ACDEFGHIKLMNPQRSTVWY"


def mutate(s, rate, rng):
    out = []
    for c in s:
        if rng.random() < rate:
            k = rng.random()
            if k < 0.05:
                continue
            elif k < 0.10:
                out.append(c)
                out.append(rng.choice(AA))
            else:
                out.append(rng.choice(AA))
        else:
            out.append(c)
    return "".join(out)


def generate(ndb, length, nq, seed, nroot=None):
    rng = random.Random(seed)
    if nroot is None:
        nroot = max(1, ndb // 20)
    roots = ["".join(rng.choice(AA) for _ in range(length)) for _ in range(nroot)]
    db = [mutate(roots[i % nroot], rng.uniform(0.05, 0.5), rng) for i in range(ndb)]
    qs = []
    for i in range(nq):
        t = rng.randrange(ndb)
        qs.append(("a%d;t=p%d" % (i, t), mutate(db[t], rng.uniform(0.05, 0.4), rng)))
    return db, qs


def main(argv):
    ndb, length, nq, seed = map(int, argv[1:5])
    pref = argv[5]
    nroot = int(argv[6]) if len(argv) > 6 else None
    db, qs = generate(ndb, length, nq, seed, nroot)
    with open(pref + ".db.fa", "w") as f:
        f.write("".join(">p%d\n%s\n" % (i, s) for i, s in enumerate(db)))
    with open(pref + ".q.fa", "w") as f:
        for lab, q in qs:
            f.write(">%s\n%s\n" % (lab, q))


if __name__ == "__main__":
    main(sys.argv)
