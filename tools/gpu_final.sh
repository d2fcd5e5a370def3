#!/bin/bash
tools/gpu_check.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1g.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_launch_r1g.log 2>&1; echo "list rc=$?"
