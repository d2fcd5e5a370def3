#!/bin/bash
# compute-sanitizer racecheck + synccheck over the small parity tests of the staged kernels, ranking and derep
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --error-exitcode 99 --print-limit 30 \
    python -m pytest tests/test_gpu_staged.py tests/test_gpu_uniques.py tests/test_gpu_stages.py -m gpu -x -q \
    -k "not more_survivors" > gpurun_out/sanitize_$tool.txt 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|Barrier error" gpurun_out/sanitize_$tool.txt | cut -c1-160 | sort | uniq -c | sort -rn | head -12
done
