#!/bin/bash
# Quick round on the B200 box: GPU tests (optionally a -k filter) + one bench run.
#   gpurun --timeout 1200 -- tools/gpu_quick.sh TAG ["pytest -k expr"]
T=${1:-q}
K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then
  timeout 1000 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_$T.txt 2>&1; echo "pytest rc=$?"
else
  timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.txt 2>&1; echo "pytest rc=$?"
fi
tail -5 gpurun_out/pytest_$T.txt
timeout 600 python bench.py --no-cpu-baseline --no-legs > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("kernels_ms_per_step"), d.get("gpu_launches"))
PY
