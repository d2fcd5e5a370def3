#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
USB_TIMING=1 tools/search_timing.sh 1000000 100000 20000 > gpurun_out/search_timing.txt 2>&1; tail -9 gpurun_out/search_timing.txt
