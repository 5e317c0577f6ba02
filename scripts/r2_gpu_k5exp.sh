#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for st in 2 3 5; do
  DMM_BUILD_DEFINES="-DK5_STAGES=$st" python -m dmm_net_b200.build --force > /dev/null 2>&1
  echo "== stages $st"; timeout 120 python scripts/prof_k5.py tc 2>&1 | grep "64 frames x 50 ROIs (typical)"
done
python -m dmm_net_b200.build --force > /dev/null 2>&1
