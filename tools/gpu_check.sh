#!/bin/bash
# Round check on the B200 box: smoke, GPU tests, both bench arms.  gpurun --timeout 1500 -- tools/gpu_check.sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_ref.json", "gpurun_out/bench.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("kernels_ms_per_step"), d.get("gpu_launches"))
PY
