#!/bin/bash
# Evidence pass for the staged candidate loop at the full workload (1 M reads x 100 k DB):
# bench line, launch list, ncu --set full of k_gate / k_dp (stage 1 launch = the big one) and k_rank.
# usage: tools/gpu_profile_stage.sh TAG [kernels: gate dp rank]
T=${1:-r2}
shift
K=${@:-gate dp}
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-legs > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("kernels_ms_per_step"), d.get("gpu_launches"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$T.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_launch_$T.log 2>&1; echo "list rc=$?"
for k in $K; do
  case $k in
    gate) ncu --set full --clock-control none --import-source on -k regex:k_gate -s 6 -c 2 -f -o gpurun_out/prof_gate_$T \
            python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_gate_$T.log 2>&1; echo "gate rc=$?";;
    dp)   ncu --set full --clock-control none --import-source on -k regex:k_dp -s 6 -c 2 -f -o gpurun_out/prof_dp_$T \
            python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_dp_$T.log 2>&1; echo "dp rc=$?";;
    rank) ncu --set full --clock-control none --import-source on -k regex:k_rank -s 3 -c 1 -f -o gpurun_out/prof_rank_$T \
            python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/ncu_rank_$T.log 2>&1; echo "rank rc=$?";;
  esac
done
