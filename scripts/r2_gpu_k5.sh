#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_container.py -x -q -k "roi_mean_pool or feature_extractor" > gpurun_out/r2_k5_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_k5_tests.txt
tail -4 gpurun_out/r2_k5_tests.txt
timeout 300 python scripts/prof_k5.py > gpurun_out/r2_k5_tc_timing.txt 2>&1
cat gpurun_out/r2_k5_tc_timing.txt
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_pool_tc_kernel -s 42 -c 1 -o gpurun_out/r2_k5_tc_full -f python scripts/prof_k5.py tc > gpurun_out/r2_k5_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_group|tc_weights|roi_pool_tc|roi_mean_pool" -s 120 -c 12 --csv --log-file gpurun_out/r2_k5_launches.csv python scripts/prof_k5.py tc > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_k5_launches.csv | cut -d, -f5,15- | head -20
fi
