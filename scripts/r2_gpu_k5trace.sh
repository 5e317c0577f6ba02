#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DMM_BUILD_DEFINES="-DDMM_TC_DEBUG" python -m dmm_net_b200.build --force > /dev/null 2>&1
timeout 120 python scripts/k5_trace.py 64 50 > gpurun_out/r2_k5_trace_64.txt 2>&1
timeout 120 python scripts/k5_trace.py 8 50 > gpurun_out/r2_k5_trace_8.txt 2>&1
python -m dmm_net_b200.build --force > /dev/null 2>&1
head -150 gpurun_out/r2_k5_trace_64.txt
