#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_exit_parity.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.txt
tail -25 gpurun_out/r2_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
