#!/bin/bash
# Full-size evidence pass (1 M reads x 100 k DB): launch list + ncu --set full of both kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_full.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_full.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rank -s 1 -c 1 -f -o gpurun_out/prof_rank_full2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rank_full2.log 2>&1; echo "rank rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -f -o gpurun_out/prof_align_full2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_align_full2.log 2>&1; echo "align rc=$?"
