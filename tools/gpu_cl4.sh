#!/bin/bash
# A/B of k_rank_big variants on one box: 1 M window-random reads, cluster_fast, USB_TIMING
mkdir -p gpurun_out
python - <<'PY'
import sys, os, time, subprocess
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import synth_np
from usearch12_b200 import build
cli = build.build_cli()
db, db_off = synth_np.gen_db(100000, 1500, seed=4)
reads, r_off, _ = synth_np.gen_reads(db, db_off, 1000000, 250, seed=3000)
synth_np.write_fasta("/tmp/r.fa", reads, r_off, "r")
for var in ("0", "1", "0", "1"):
    t = time.time()
    r = subprocess.run([cli, "-cluster_fast", "/tmp/r.fa", "-id", "0.97", "-uc", "/tmp/o%s.uc" % var],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, USB_TIMING="1", USB_BIG_VARIANT=var))
    print("variant", var, "%.1fs" % (time.time() - t), r.returncode, [l for l in r.stdout.splitlines() if "big path" in l])
print("same uc:", open("/tmp/o0.uc", "rb").read() == open("/tmp/o1.uc", "rb").read())
PY
USB_TIMING=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-legs 2>&1 | tail -4
