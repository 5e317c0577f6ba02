#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DMM_BUILD_DEFINES="-DDMM_TC_DEBUG" python -m dmm_net_b200.build --force > /dev/null 2>&1
for seed in 1 2 3; do
timeout 200 python scripts/k5_stress.py $seed 8000 2>&1 | grep -v "^  File\|^    \|Warning" | sort | uniq -c | sort -rn | head -14
done
python -m dmm_net_b200.build --force > /dev/null 2>&1
