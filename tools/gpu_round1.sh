#!/bin/bash
# First measurement pass on the B200 box: smoke, bench (config 2 + full), ncu launch list + full captures.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --reads 100000 --db 10000 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv \
  python bench.py --reads 100000 --db 10000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -f -o gpurun_out/prof_align \
  python bench.py --reads 100000 --db 10000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_align.log 2>&1; echo "ncu align rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rank -s 1 -c 1 -f -o gpurun_out/prof_rank \
  python bench.py --reads 100000 --db 10000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rank.log 2>&1; echo "ncu rank rc=$?"
tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
nproc; free -g | head -2
