#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_exit_parity.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.txt
tail -6 gpurun_out/r2_gputests.txt
python scripts/prof_k5.py tc 2>&1 | grep "bwd\|64 frames x 50 ROIs (typical)"
