#!/bin/bash
mkdir -p gpurun_out
tools/search_timing.sh 1000000 100000 20000 > gpurun_out/search_timing.txt 2>&1; tail -12 gpurun_out/search_timing.txt
