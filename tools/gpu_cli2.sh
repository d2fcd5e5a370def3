#!/bin/bash
# host CLI with one and with two devices: 1 M reads vs 100 k targets, FASTA in -> .uc + .b6 out
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import synth_np
from usearch12_b200 import build
build.build_cli()
db, db_off = synth_np.gen_db(100000, 1500, seed=4)
reads, r_off, _ = synth_np.gen_reads(db, db_off, 1000000, 250, seed=1000)
synth_np.write_fasta("/tmp/db.fa", db, db_off, "db")
synth_np.write_fasta("/tmp/q.fa", reads, r_off, "q")
PY
for g in 1 2 1 2; do
  /usr/bin/env USB_TIMING=1 usearch12_b200/usearch12_b200_cli -usearch_global /tmp/q.fa -db /tmp/db.fa -id 0.97 -strand plus -gpus $g \
    -uc /tmp/o$g.uc -blast6out /tmp/o$g.b6 2>&1 | grep -E "timing|GPU" | sed "s/^/gpus=$g /"
done
cmp /tmp/o1.uc /tmp/o2.uc && cmp /tmp/o1.b6 /tmp/o2.b6 && echo "outputs identical for 1 and 2 GPUs"
