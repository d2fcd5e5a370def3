#!/bin/bash
# racecheck after the __syncwarp fixes: one staged test (k_gate, k_dp, k_commit and the one-kernel k_align)
mkdir -p gpurun_out
timeout 540 compute-sanitizer --tool racecheck --target-processes all --print-limit 30 \
  python -m pytest tests/test_gpu_staged.py -m gpu -x -q -k "many_stages" > gpurun_out/sanitize_racecheck2.txt 2>&1
echo "racecheck rc=$?"
grep -E "Race reported|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck2.txt | cut -c1-200 | sort | uniq -c | sort -rn | head -12
