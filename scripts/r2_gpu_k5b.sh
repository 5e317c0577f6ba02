#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_container.py -x -q -k "roi_mean_pool or feature_extractor" > gpurun_out/r2_k5_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_k5_tests.txt
tail -4 gpurun_out/r2_k5_tests.txt
timeout 300 python scripts/prof_k5.py tc > gpurun_out/r2_k5_tc_timing.txt 2>&1
cat gpurun_out/r2_k5_tc_timing.txt
DMM_BUILD_DEFINES="-DDMM_TC_DEBUG" python -m dmm_net_b200.build --force > /dev/null 2>&1
timeout 120 python scripts/k5_trace.py 64 50 > gpurun_out/r2_k5_trace_64.txt 2>&1
timeout 120 python scripts/k5_trace.py 8 50 > gpurun_out/r2_k5_trace_8.txt 2>&1
python -m dmm_net_b200.build --force > /dev/null 2>&1
head -120 gpurun_out/r2_k5_trace_64.txt
