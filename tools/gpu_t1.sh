#!/bin/bash
# all GPU tests, then the default bench (no legs) with the phase timers of usb_search_batch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
USB_TIMING=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/bench_t1.json 2> gpurun_out/bench_t1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_t1.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d["e2e"], d.get("kernels_ms_per_step"), d.get("gpu_launches"))
PY
grep usb_search_batch gpurun_out/bench_t1.err
