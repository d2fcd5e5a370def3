#!/bin/bash
# cluster_fast check on the B200 box: parity tests, then the two config-3 legs of bench.py.
T=${1:-cl}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cluster" > gpurun_out/pytest_$T.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$T.txt
timeout 1500 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --legs cluster,cluster_amplicon > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
for k, v in (d.get("legs") or {}).items():
    print("leg", k, {x: (round(y, 1) if isinstance(y, float) else y) for x, y in v.items() if x not in ("workload", "what")})
PY
