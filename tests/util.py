"""Test helpers: golden fixture access, FASTA reading, and the reference's output line formats
(re-stated here only to compare product hits with the golden files; formats follow
userout.cpp:150-215, blast6out.cpp:27-80, outputuc.cpp:19-69)."""
import gzip
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

VARIANTS = {
    "plus97": dict(id=0.97, strand_both=0),
    "both97": dict(id=0.97, strand_both=1),
    "plus90_ma4": dict(id=0.9, strand_both=0, maxaccepts=4, maxrejects=64),
    "both80_ma0": dict(id=0.8, strand_both=1, maxaccepts=3, maxrejects=16),
}


def read_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    labels, seqs, cur = [], [], None
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if cur is not None and cur[1]:
                    labels.append(cur[0])
                    seqs.append("".join(cur[1]))
                cur = (line[1:], [])
            elif cur is not None:
                cur[1].append("".join(c for c in line if c.isalpha()))
    if cur is not None and cur[1]:
        labels.append(cur[0])
        seqs.append("".join(cur[1]))
    seqs = [s for s in seqs]
    return labels, seqs


class Golden:
    def __init__(self):
        self.db_labels, self.db = read_fasta(os.path.join(GOLDEN, "db.fa.gz"))
        self.q_labels, self.q = read_fasta(os.path.join(GOLDEN, "q.fa.gz"))

    def lines(self, variant, kind):
        with gzip.open(os.path.join(GOLDEN, "%s.%s.gz" % (variant, kind)), "rt") as f:
            return f.read().splitlines()


def pct(ids, alnlen):
    return 100.0 * (float(ids) / float(alnlen) if alnlen else 0.0)


def fmt_user(h, cigar, ql, tl, nucleo=True):
    return "%s\t%s\t%.1f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%s\t%s" % (
        ql, tl, pct(h["ids"], h["alnlen"]), h["alnlen"], h["mism"], h["opens"], 1, h["ql"], 1, h["tl"], cigar,
        ("-" if h["strand"] else "+") if nucleo else ".")


def fmt_b6(h, ql, tl):
    tlo, thi = (h["tl"], 1) if h["strand"] else (1, h["tl"])
    return "%s\t%s\t%.1f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t*\t*" % (
        ql, tl, pct(h["ids"], h["alnlen"]), h["alnlen"], h["mism"], h["opens"], 1, h["ql"], tlo, thi)


def fmt_uc_hit(h, cigar, ql, tl, nucleo=True):
    return "H\t%d\t%d\t%.1f\t%s\t0\t0\t%s\t%s\t%s" % (
        h["target"], h["ql"], pct(h["ids"], h["alnlen"]), ("-" if h["strand"] else "+") if nucleo else ".", cigar, ql, tl)


def fmt_uc_nohit(qlen, ql):
    return "N\t*\t%d\t*\t.\t*\t*\t*\t%s\t*" % (qlen, ql)


def product_lines(res, q_labels, q_seqs, db_labels, nucleo=True):
    """usb200 Result -> (user, uc, b6) line lists in query input order."""
    user, uc, b6 = [], [], []
    for qi in range(len(q_seqs)):
        b, e = int(res.qoff[qi]), int(res.qoff[qi + 1])
        for k in range(b, e):
            h = res.hits[k]
            cig = res.cigar(h)
            tl = db_labels[int(h["target"])]
            user.append(fmt_user(h, cig, q_labels[qi], tl, nucleo))
            uc.append(fmt_uc_hit(h, cig, q_labels[qi], tl, nucleo))
            b6.append(fmt_b6(h, q_labels[qi], tl))
        if b == e:
            uc.append(fmt_uc_nohit(len(q_seqs[qi]), q_labels[qi]))
    return user, uc, b6


def oracle_lines(searcher, q_labels, q_seqs, db_labels, nucleo=True):
    from oracle import uso_py as O
    user, uc, b6 = [], [], []
    for qi, s in enumerate(q_seqs):
        hits = searcher.search(s, qi)
        for h in hits:
            cig = O.compress_path(h["path"])
            tl = db_labels[h["target"]]
            user.append(fmt_user(h, cig, q_labels[qi], tl, nucleo))
            uc.append(fmt_uc_hit(h, cig, q_labels[qi], tl, nucleo))
            b6.append(fmt_b6(h, q_labels[qi], tl))
        if not hits:
            uc.append(fmt_uc_nohit(len(s), q_labels[qi]))
    return user, uc, b6


def first_diff(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return "line %d:\n  got  %s\n  want %s" % (i, x, y)
    if len(a) != len(b):
        return "length %d vs %d" % (len(a), len(b))
    return None


def mutate(s, rate, rng, alphabet="ACGT"):
    out = []
    for c in s:
        if rng.random() < rate:
            k = rng.random()
            if k < 0.1:
                continue
            elif k < 0.2:
                out.append(c)
                out.append(rng.choice(alphabet))
            else:
                out.append(rng.choice([x for x in alphabet if x != c]))
        else:
            out.append(c)
    return "".join(out)


# ---------------------------------------------------------------- amino acid usearch_global
# tools/make_golden_aa_global.py: config 1 (the reference's tmp/test.fa vs itself) and the protein
# families of the usearch_local fixtures searched globally
AA_GLOBAL_VARIANTS = {
    "cfg1_id90": dict(inputs="cfg1", id=0.9),
    "cfg1_id30_ma8": dict(inputs="cfg1", id=0.3, maxaccepts=8, maxrejects=64),
    "gaa_id50": dict(inputs="gaa", id=0.5),
    "gaa_id30_ma8": dict(inputs="gaa", id=0.3, maxaccepts=8, maxrejects=64),
    "gaa_id90": dict(inputs="gaa", id=0.9),
    "gaa_id70_ma3": dict(inputs="gaa", id=0.7, maxaccepts=3, maxrejects=16),
}


def aa_global_inputs(kind):
    """-> (db_labels, db, q_labels, q)"""
    if kind == "cfg1":
        labels, seqs = read_fasta(os.path.join(GOLDEN, "cfg1_test.fa.gz"))
        return labels, seqs, labels, seqs
    dl, d = read_fasta(os.path.join(GOLDEN, "loc_aa_db.fa.gz"))
    ql, q = read_fasta(os.path.join(GOLDEN, "loc_aa_q.fa.gz"))
    return dl, d, ql, q


def golden_lines(variant, kind):
    with gzip.open(os.path.join(GOLDEN, "%s.%s.gz" % (variant, kind)), "rt") as f:
        return f.read().splitlines()


# ---------------------------------------------------------------- usearch_local
LOCAL_VARIANTS = {
    "loc_aa_e5": dict(nucleo=False, id=0.5, evalue=1e-5),
    "loc_aa_ma4": dict(nucleo=False, id=0.3, evalue=10.0, maxaccepts=4, maxrejects=64),
    "loc_nt_plus": dict(nucleo=True, id=0.9, evalue=1e-5, maxaccepts=2, maxrejects=16),
    "loc_nt_both": dict(nucleo=True, id=0.8, evalue=1e-3, maxaccepts=3, maxrejects=8, strand_both=1),
}


# sequences above g_MaxL = 4096 letters: split X-drop extensions (tools/make_golden_local_long.py)
LOCAL_LONG_VARIANTS = {
    "loclong_nt_both": dict(nucleo=True, id=0.7, evalue=1e-5, maxaccepts=2, maxrejects=16, strand_both=1),
    "loclong_aa_e5": dict(nucleo=False, id=0.5, evalue=1e-5),
}


class GoldenLocal:
    def __init__(self, kind, prefix="loc"):
        self.db_labels, self.db = read_fasta(os.path.join(GOLDEN, "%s_%s_db.fa.gz" % (prefix, kind)))
        self.q_labels, self.q = read_fasta(os.path.join(GOLDEN, "%s_%s_q.fa.gz" % (prefix, kind)))

    def lines(self, variant, kind):
        with gzip.open(os.path.join(GOLDEN, "%s.%s.gz" % (variant, kind)), "rt") as f:
            return f.read().splitlines()


def _strand_char(h, nucleo):
    return ("-" if h["strand"] else "+") if nucleo else "."


def fmt_local(h, cigar, evalue, bits, ql, tl, nucleo):
    """(user, uc, b6) lines of one local hit; -userfields
    query+target+id+alnlen+mism+opens+qlo+qhi+tlo+thi+evalue+bits+raw+caln+qstrand (userout.cpp:150-215),
    blast6out.cpp:27-80, outputuc.cpp:45-69.  h holds loi/hii/loj/hij (0-based segment ends) and raw."""
    p = pct(h["ids"], h["alnlen"])
    st = _strand_char(h, nucleo)
    # query coordinates are reported on the plus strand (arscorer.cpp:683-745); blast6 swaps the
    # target ends for a reverse-complemented query (arscorer.cpp:748-808)
    rc = bool(h["strand"])
    qlo = h["ql"] - h["hii"] - 1 if rc else h["loi"]
    qhi = h["ql"] - h["loi"] - 1 if rc else h["hii"]
    tlo, thi = h["loj"], h["hij"]
    user = "%s\t%s\t%.1f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.3g\t%.0f\t%.0f\t%s\t%s" % (
        ql, tl, p, h["alnlen"], h["mism"], h["opens"], qlo + 1, qhi + 1, tlo + 1, thi + 1, evalue, bits,
        h["raw"], cigar, st)
    uc = "H\t%d\t%d\t%.1f\t%s\t%d\t%d\t%s\t%s\t%s" % (h["target"], h["ql"], p, st, qlo, tlo, cigar, ql, tl)
    b6 = "%s\t%s\t%.1f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2g\t%.1f" % (
        ql, tl, p, h["alnlen"], h["mism"], h["opens"], qlo + 1, qhi + 1, thi + 1 if rc else tlo + 1,
        tlo + 1 if rc else thi + 1, evalue, bits)
    return user, uc, b6


def product_lines_local(res, searcher, q_labels, q_seqs, db_labels, nucleo):
    user, uc, b6 = [], [], []
    for qi in range(len(q_seqs)):
        b, e = int(res.qoff[qi]), int(res.qoff[qi + 1])
        for k in range(b, e):
            h = res.hits[k]
            d = dict(ids=int(h["ids"]), alnlen=int(h["alnlen"]), mism=int(h["mism"]), opens=int(h["opens"]),
                     loi=int(h["first_mq"]), hii=int(h["last_mq"]), loj=int(h["first_mt"]), hij=int(h["last_mt"]),
                     raw=int(h["raw"]), target=int(h["target"]), ql=int(h["ql"]), strand=int(h["strand"]))
            ev, bits = searcher.local_evalue(d["raw"], d["ql"])
            u, c, x = fmt_local(d, res.cigar(h), ev, bits, q_labels[qi], db_labels[d["target"]], nucleo)
            user.append(u)
            uc.append(c)
            b6.append(x)
        if b == e:
            uc.append(fmt_uc_nohit(len(q_seqs[qi]), q_labels[qi]))
    return user, uc, b6


def oracle_lines_local(searcher, q_labels, q_seqs, db_labels, nucleo):
    from oracle import uso_py as O
    user, uc, b6 = [], [], []
    for qi, s in enumerate(q_seqs):
        hits = searcher.search(s, qi)
        for h in hits:
            d = dict(h)
            d["hii"] = h["loi"] + h["leni"] - 1
            d["hij"] = h["loj"] + h["lenj"] - 1
            u, c, x = fmt_local(d, O.compress_path(h["path"]), h["evalue"], h["bits"], q_labels[qi], db_labels[h["target"]],
                                nucleo)
            user.append(u)
            uc.append(c)
            b6.append(x)
        if not hits:
            uc.append(fmt_uc_nohit(len(s), q_labels[qi]))
    return user, uc, b6


def oracle_local_params(nucleo, **kw):
    from oracle import uso_py as O
    return O.default_params(amino=not nucleo, local=1, **kw)


def product_local_params(nucleo, evalue, **kw):
    import ctypes as C
    from usearch12_b200 import capi
    p = capi.default_params(**kw)
    capi.lib().usb_set_local(C.byref(p), int(nucleo), float(evalue))
    return p


# -cluster_fast with size annotations: option sets of the golden files cluster_<name>.* (made by
# tools/make_golden_cluster_sizes.py with the reference binary)
CLUSTER_SIZE_VARIANTS = {
    "szin": ["-sizein", "-sizeout", "-sort", "size"],
    "szout": ["-sizeout", "-relabel", "Otu", "-minsize", "3", "-sort", "size"],
    "szin_len": ["-sizein", "-sort", "length"],
}


def sized_cluster_reads(path):
    """tests/golden/cluster_reads.fa.gz with ';size=N;' appended to every label."""
    import gzip
    labels, seqs = read_fasta(os.path.join(GOLDEN, "cluster_reads.fa.gz"))
    with open(path, "w") as f:
        for i, (lab, s) in enumerate(zip(labels, seqs)):
            f.write(">%s;size=%d;\n%s\n" % (lab, 1 + (i * 7) % 13, s if isinstance(s, str) else s.decode()))


# -cluster_smallmem golden variants (tools/make_golden_cluster_sizes.py): name -> (input order, options)
SMALLMEM_VARIANTS = {
    "sm_len": ("length", []),
    "sm_size": ("size", ["-sortedby", "size", "-sizein", "-sizeout", "-minsize", "4"]),
    "sm_other": ("none", ["-sortedby", "other", "-maxrejects", "8"]),
}


def smallmem_reads(path, order):
    """cluster_reads.fa.gz with size annotations, sorted for -cluster_smallmem (stable): by decreasing
    length, by decreasing size annotation, or in file order."""
    labels, seqs = read_fasta(os.path.join(GOLDEN, "cluster_reads.fa.gz"))
    recs = [("%s;size=%d;" % (lab, 1 + (i * 7) % 13), s if isinstance(s, str) else s.decode(), 1 + (i * 7) % 13)
            for i, (lab, s) in enumerate(zip(labels, seqs))]
    if order == "length":
        recs.sort(key=lambda r: -len(r[1]))
    elif order == "size":
        recs.sort(key=lambda r: -r[2])
    with open(path, "w") as f:
        for lab, s, _ in recs:
            f.write(">%s\n%s\n" % (lab, s))


# -fastx_uniques with -sizein / -topn: option sets of tests/golden/uniq2_<name>.fa.gz
# (tools/make_golden_uniques.py, reference binary)
UNIQUES2_VARIANTS = {
    "sizein": ["-sizein", "-sizeout", "-relabel", "U"],
    "topn": ["-sizeout", "-topn", "40", "-minuniquesize", "2"],
}
