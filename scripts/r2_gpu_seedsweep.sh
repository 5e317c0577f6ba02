#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DMM_BUILD_DEFINES="-DDMM_TC_DEBUG" python -m dmm_net_b200.build --force > /dev/null 2>&1
timeout 600 python scripts/train_seed_sweep.py 3 0 1 2 4 5 6 7 2>&1 | grep -v "^  File\|^    \|Warning" | sort | uniq -c | sort -rn | head -30
python -m dmm_net_b200.build --force > /dev/null 2>&1
