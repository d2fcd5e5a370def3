#!/bin/bash
# tools/search_timing.sh READS DB [REF_SAMPLE]: whole-command timing of the -usearch_global CLI
# (FASTA parse, index build + upload, search, .uc/.b6 writing) next to the reference binary on a sample.
python - "$@" <<PY
import sys, os, time, subprocess
sys.path.insert(0,"tools"); sys.path.insert(0,".")
import synth_np
from usearch12_b200 import build
cli = build.build_cli()
n, d = int(sys.argv[1]), int(sys.argv[2])
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
db, db_off = synth_np.gen_db(d, 1500, seed=4)
reads, r_off, _ = synth_np.gen_reads(db, db_off, n, 250, seed=1000)
synth_np.write_fasta("/tmp/db.fa", db, db_off, "db")
synth_np.write_fasta("/tmp/q.fa", reads, r_off, "q")
synth_np.write_fasta("/tmp/qs.fa", reads, r_off, "q", 0, ns)
for k in range(2):
    t=time.time()
    r=subprocess.run([cli,"-usearch_global","/tmp/q.fa","-db","/tmp/db.fa","-id","0.97","-strand","plus","-uc","/tmp/o.uc","-blast6out","/tmp/o.b6"],
                     env=dict(os.environ, USB_TIMING="1"),stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True)
    dt=time.time()-t
    print(r.stdout[-800:]); print("usb200 CLI run %d: total %.2fs  %.0f reads/s (whole command, %d reads x %d targets)" % (k, dt, n/dt, n, d))
ref="oracle/_ref/usearch12"
if os.path.exists(ref):
    t=time.time(); subprocess.run([ref,"-usearch_global","/tmp/qs.fa","-db","/tmp/db.fa","-id","0.97","-strand","plus","-uc","/tmp/r.uc","-blast6out","/tmp/r.b6","-threads",str(os.cpu_count()),"-quiet"],check=True,stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL); dt=time.time()-t
    print("reference CLI: %d reads in %.2fs whole command (%d threads) = %.0f reads/s" % (ns, dt, os.cpu_count(), ns/dt))
    # the reference's .uc lines (any order) must equal ours for the sampled queries (labels q0 .. q<ns-1>)
    want = sorted(open("/tmp/r.uc").read().splitlines())
    got = sorted(l for l in open("/tmp/o.uc").read().splitlines() if int(l.split("\t")[8][1:]) < ns)
    print("uc lines of the %d sampled queries identical to the reference binary: %s (%d lines)" % (ns, want == got, len(want)))
    want = sorted(open("/tmp/r.b6").read().splitlines())
    got = sorted(l for l in open("/tmp/o.b6").read().splitlines() if int(l.split("\t")[0][1:]) < ns)
    print("blast6 lines identical: %s (%d lines)" % (want == got, len(want)))
PY
