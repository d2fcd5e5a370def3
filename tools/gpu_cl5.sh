#!/bin/bash
# 4-bit counters in k_rank_big: parity tests of the big path, then A/B (USB_BIG_VARIANT=2 forces byte counters)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "big or cluster or rank" 2>&1 | tail -4
python - <<'PY'
import sys, os, time, subprocess
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import synth_np
from usearch12_b200 import build
cli = build.build_cli()
db, db_off = synth_np.gen_db(100000, 1500, seed=4)
reads, r_off, _ = synth_np.gen_reads(db, db_off, 1000000, 250, seed=3000)
synth_np.write_fasta("/tmp/r.fa", reads, r_off, "r")
for var in ("0", "2", "0", "2", "0", "2"):
    t = time.time()
    r = subprocess.run([cli, "-cluster_fast", "/tmp/r.fa", "-id", "0.97", "-uc", "/tmp/o%s.uc" % var],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, USB_TIMING="1", USB_BIG_VARIANT=var))
    print("variant", var, "%.1fs" % (time.time() - t), r.returncode, [l.split("kernels")[1] for l in r.stdout.splitlines() if "big path" in l])
print("same uc:", open("/tmp/o0.uc", "rb").read() == open("/tmp/o2.uc", "rb").read())
PY
