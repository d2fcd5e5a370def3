#!/bin/bash
# Full-size evidence pass (1 M reads x 100 k DB): launch list + ncu --set full of both kernels.
# usage: tools/gpu_profile_full.sh TAG   -> gpurun_out/launches_TAG.csv, prof_rank_TAG.ncu-rep, prof_align_TAG.ncu-rep
T=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$T.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_launch_$T.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rank -s 1 -c 1 -f -o gpurun_out/prof_rank_$T \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_rank_$T.log 2>&1; echo "rank rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -f -o gpurun_out/prof_align_$T \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_align_$T.log 2>&1; echo "align rc=$?"
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_$T.json
