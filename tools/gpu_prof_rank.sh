#!/bin/bash
# ncu --set full capture of k_rank at the full workload (1 M reads x 100 k DB)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_rank -s 1 -c 1 -f -o gpurun_out/prof_rank_${1:-r1d} \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rank.log 2>&1; echo "rank rc=$?"
ls -la gpurun_out/*.ncu-rep
