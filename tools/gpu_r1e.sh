#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; rc=$?
  python - "$name" $rc <<'PY'
import json,sys
f="gpurun_out/bench_%s.json"%sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(sys.argv[1], "k_rank %.1f ms"%d["kernels_ms_per_step"]["k_rank"], "GB/s %.0f"%d["roofline"]["other"]["k_rank_GBps"], "hits", d["config"]["hit_rate"], "e2e", round(d["e2e"]["value"]), "value", round(d["value"]), "ix_s", d["config"]["index_build_s"])
except Exception as e:
    print(sys.argv[1], "FAILED rc", sys.argv[2], e)
PY
  grep "phase cycles" gpurun_out/bench_$name.err | tail -1
}
run e64_p USB_RANK_PROF=1
run e0_p USB_RANK_EARLY=0 USB_RANK_PROF=1
run e128_p USB_RANK_EARLY=128 USB_RANK_PROF=1
