#!/bin/bash
# big-database ranking + cluster_fast check: parity tests, then the two 1 M-read cluster runs with timers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "big or cluster or rank or otutab" 2>&1 | tail -5
python - <<'PY'
import sys, os, time, subprocess
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import synth_np
from usearch12_b200 import build
cli = build.build_cli()
db, db_off = synth_np.gen_db(100000, 1500, seed=4)
for amp in (False, True):
    reads, r_off, _ = synth_np.gen_reads(db, db_off, 1000000, 250, seed=3000, window=(500, 750) if amp else None)
    synth_np.write_fasta("/tmp/r.fa", reads, r_off, "r")
    for rep in range(2):
        t = time.time()
        r = subprocess.run([cli, "-cluster_fast", "/tmp/r.fa", "-id", "0.97", "-uc", "/tmp/o.uc", "-centroids", "/tmp/o.fa"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, USB_TIMING="1"))
        print("amplicon" if amp else "window-random", "%.1fs" % (time.time() - t), r.returncode)
    print(r.stdout[-1800:])
PY
