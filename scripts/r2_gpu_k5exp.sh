#!/bin/bash
cd "$(dirname "$0")/.."
for st in 2 4; do
  DMM_BUILD_DEFINES="-DK5_STAGES=$st" python -m dmm_net_b200.build --force > /dev/null 2>&1
  echo "stages $st: $(python scripts/prof_k5.py tc 2>&1 | grep '64 frames x 50 ROIs (typical)\|64 frames x 50 ROIs (whole' | cut -d: -f2 | tr '\n' ' ')"
done
python -m dmm_net_b200.build --force > /dev/null 2>&1
