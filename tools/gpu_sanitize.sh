#!/bin/bash
# compute-sanitizer memcheck over the small parity tests (kernels under the sanitizer run 10-50x slower)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_gpu_stages.py tests/test_gpu_uniques.py tests/test_gpu_staged.py tests/test_gpu_accept.py -m gpu -x -q \
  -k "not fullsize and not 120k" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitize_memcheck.txt | sort | uniq -c | sort -rn | head -20
