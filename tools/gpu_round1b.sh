#!/bin/bash
# Round-1 second evidence pass: whole GPU test suite, default bench line, config-5 local bench,
# launch list + ncu --set full of k_local.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cat gpurun_out/bench_full.json
timeout 600 python tools/bench_local.py > gpurun_out/bench_local.json 2> gpurun_out/bench_local.err; echo "bench_local rc=$?"; cat gpurun_out/bench_local.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_local.csv \
  python tools/bench_local.py --steps 1 --check 0 --ref-sample 0 > gpurun_out/ncu_launch_local.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_local -s 1 -c 1 -f -o gpurun_out/prof_local \
  python tools/bench_local.py --steps 1 --check 0 --ref-sample 0 > gpurun_out/ncu_local.log 2>&1; echo "local rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rank -s 1 -c 1 -f -o gpurun_out/prof_rank_local \
  python tools/bench_local.py --steps 1 --check 0 --ref-sample 0 > gpurun_out/ncu_rank_local.log 2>&1; echo "rank rc=$?"
