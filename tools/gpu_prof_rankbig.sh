#!/bin/bash
# ncu --set full of one late k_rank_big launch of a 1 M-read cluster_fast run (window-random reads)
T=${1:-r2}
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import synth_np
db, db_off = synth_np.gen_db(100000, 1500, seed=4)
reads, r_off, _ = synth_np.gen_reads(db, db_off, 1000000, 250, seed=3000)
synth_np.write_fasta("/tmp/r.fa", reads, r_off, "r")
PY
CLI=usearch12_b200/usearch12_b200_cli
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rank_big -s 40 -c 1 -f -o gpurun_out/prof_rankbig_$T \
  $CLI -cluster_fast /tmp/r.fa -id 0.97 -uc /tmp/o.uc > gpurun_out/ncu_rankbig_$T.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_rankbig_$T.log
