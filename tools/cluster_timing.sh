#!/bin/bash
# tools/cluster_timing.sh READS DB : cluster_fast CLI timing with the round timers (USB_TIMING=1)
python - "$@" <<PY
import sys, os, time, subprocess
sys.path.insert(0,"tools"); sys.path.insert(0,".")
import synth_np
from usearch12_b200 import build
cli = build.build_cli()
n, d = int(sys.argv[1]), int(sys.argv[2])
amp = len(sys.argv) > 3
db, db_off = synth_np.gen_db(d, 1500, seed=4)
reads, r_off, _ = synth_np.gen_reads(db, db_off, n, 250, seed=3000, window=(500,750) if amp else None)
synth_np.write_fasta("/tmp/r.fa", reads, r_off, "r")
t=time.time()
env=dict(os.environ, USB_TIMING="1")
r=subprocess.run([cli,"-cluster_fast","/tmp/r.fa","-id","0.97","-uc","/tmp/o.uc"],env=env,stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True)
print(r.stdout[-600:]); print("total %.2fs  %.0f reads/s" % (time.time()-t, n/(time.time()-t)))
PY
