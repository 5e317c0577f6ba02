#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DMM_BUILD_DEFINES="-DDMM_TC_DEBUG" python -m dmm_net_b200.build --force > gpurun_out/r2_k5dbg_build.txt 2>&1 || tail -20 gpurun_out/r2_k5dbg_build.txt
timeout 120 python scripts/k5_debug.py 1 5 2>&1 | grep -v "^  File\|^    " | sort | uniq -c | sort -rn | head -12
timeout 120 python scripts/k5_debug.py 8 50 2>&1 | grep -v "^  File\|^    " | sort | uniq -c | sort -rn | head -12
python -m dmm_net_b200.build --force > /dev/null 2>&1
bash scripts/r2_gpu_k5.sh
