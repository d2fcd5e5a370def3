#!/bin/bash
# k_rank phase-locked row order: parity under the option, then a sweep of the period.
mkdir -p gpurun_out
USB_RANK_PERIOD_NS=60000 timeout 900 python -m pytest tests -m gpu -x -q -k "search or stages or staged or fullsize" > gpurun_out/pytest_period.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_period.txt
for ns in 0 20000 40000 60000 80000 120000 200000; do
  USB_RANK_PERIOD_NS=$ns python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/bench_per_$ns.json 2> gpurun_out/bench_per_$ns.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_per_$ns.json").read().strip().splitlines()[-1])
print("period $ns ns:", {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items() if k.startswith("k_")}, "value %.0f" % d["value"])
PY
done
