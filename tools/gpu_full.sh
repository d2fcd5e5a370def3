#!/bin/bash
# Whole round check on the B200 box: GPU tests, the default bench line (cpu baseline + legs).
#   gpurun --timeout 2400 -- tools/gpu_full.sh TAG
T=${1:-full}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$T.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
S0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$? in $(( $(date +%s) - S0 )) s"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("kernels_ms_per_step"), d.get("gpu_launches"))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline"))
for k, v in (d.get("legs") or {}).items():
    print("leg", k, {x: (round(y, 1) if isinstance(y, float) else y) for x, y in v.items() if x not in ("workload", "what")})
PY
S0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "reference arm rc=$? in $(( $(date +%s) - S0 )) s"
tail -c 600 gpurun_out/bench_ref_$T.json
