#!/bin/bash
# all GPU tests, then the cluster legs, then the main bench line
T=${1:-mix}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.txt
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --legs cluster,cluster_amplicon > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items() if k.startswith("k_")})
for k, v in (d.get("legs") or {}).items():
    print("leg", k, {x: (round(y, 1) if isinstance(y, float) else y) for x, y in v.items() if x not in ("workload", "what")})
PY
