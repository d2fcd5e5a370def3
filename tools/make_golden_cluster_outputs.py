#!/usr/bin/env python3
"""Golden files for the per-hit output files of -cluster_fast (the OutputSink that MakeClusterSearcher puts next to
the ClusterSink): tests/golden/cluster_out.{user,b6}.gz and digests of -alnout (without its two header lines),
-fastapairs, -matched, -notmatched in cluster_out_sha256.json, from the UNMODIFIED reference binary (-threads 1) on
tests/golden/cluster_reads.fa.gz.   Usage: python tools/make_golden_cluster_outputs.py"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "usearch12")
G = os.path.join(ROOT, "tests", "golden")
USERFIELDS = "query+target+id+clusternr+caln+qstrand+ql+tl+qrow"
OPTS = ["-id", "0.97", "-sort", "length"]
FLAGS = {"user": "-userout", "b6": "-blast6out", "aln": "-alnout", "pairs": "-fastapairs", "matched": "-matched",
         "notmatched": "-notmatched", "uc": "-uc"}


def digest_of(kind, data):
    if kind == "aln":
        data = b"\n".join(data.split(b"\n")[2:])
    return {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}


def main():
    with tempfile.TemporaryDirectory() as tmp:
        reads = os.path.join(tmp, "r.fa")
        with gzip.open(os.path.join(G, "cluster_reads.fa.gz"), "rb") as f, open(reads, "wb") as g:
            g.write(f.read())
        outs = {k: os.path.join(tmp, "o." + k) for k in FLAGS}
        cmd = [REF, "-cluster_fast", reads, "-threads", "1", "-quiet", "-userfields", USERFIELDS] + OPTS
        for k, flag in FLAGS.items():
            cmd += [flag, outs[k]]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with gzip.open(os.path.join(G, "cluster_length.uc.gz"), "rb") as f:  # the .uc is the one of the -sort length golden
            assert f.read() == open(outs["uc"], "rb").read()
        sums = {}
        for k in FLAGS:
            data = open(outs[k], "rb").read()
            if k in ("user", "b6"):
                with gzip.GzipFile(os.path.join(G, "cluster_out.%s.gz" % k), "wb", compresslevel=9, mtime=0) as g:
                    g.write(data)
                print("cluster_out", k, data.count(b"\n"), "lines")
            elif k != "uc":
                sums[k] = digest_of(k, data)
        with open(os.path.join(G, "cluster_out_sha256.json"), "w") as f:
            json.dump(sums, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    sys.exit(main())
