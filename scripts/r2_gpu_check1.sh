#!/bin/bash
# round 2, GPU call 1: gpu tests, train-step phase timing on 1 GPU, K5 baseline timing + ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_exit_parity.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.txt
tail -5 gpurun_out/r2_gputests.txt
python examples/synthetic_train_step.py --steps 8 --warmup 3 --reduce none > gpurun_out/r2_train_n1.txt 2>&1
cat gpurun_out/r2_train_n1.txt | tail -3
python scripts/prof_k5.py > gpurun_out/r2_k5_baseline.txt 2>&1
cat gpurun_out/r2_k5_baseline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_mean_pool -s 6 -c 1 -o gpurun_out/r2_k5_simt_full -f python scripts/prof_k5.py > gpurun_out/r2_k5_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
