#!/bin/bash
# all GPU tests, then the default bench with every leg (no cpu baselines)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
USB_TIMING=1 timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_t2.json 2> gpurun_out/bench_t2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_t2.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("kernels_ms_per_step"), d.get("gpu_launches"), d["roofline"]["traffic"])
for k, v in (d.get("legs") or {}).items():
    print("leg", k, {x: (round(y, 1) if isinstance(y, float) else y) for x, y in v.items() if x not in ("workload", "what")})
PY
grep "usb_search_batch" gpurun_out/bench_t2.err | tail -2
