#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_exit_parity.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.txt
tail -15 gpurun_out/r2_gputests.txt
python scripts/prof_k5.py tc > gpurun_out/r2_k5_tc_timing.txt 2>&1; cat gpurun_out/r2_k5_tc_timing.txt
python scripts/prof_k1_traffic.py 256 train; python scripts/prof_k1_traffic.py 256
date +%s > gpurun_out/t0; python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - $(cat gpurun_out/t0) )) s"
grep -v Warning gpurun_out/r2_bench_n1.err | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_tc_kernel -s 42 -c 1 -o gpurun_out/r2_k5_tc_full -f python scripts/prof_k5.py tc > gpurun_out/r2_k5_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_kernel -s 3 -c 1 -o gpurun_out/r2_k5_bwd_full -f python scripts/prof_k5.py tc > gpurun_out/r2_k5_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_group|tc_weights|roi_pool" -s 130 -c 12 --csv --log-file gpurun_out/r2_k5_launches.csv python scripts/prof_k5.py tc > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:mask_iou_partial_tma -s 3 -c 1 -o gpurun_out/r2_k1_eval_full -f python scripts/prof_k1_traffic.py 1024 > gpurun_out/r2_k1_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:mask_iou_partial_tma -s 3 -c 1 -o gpurun_out/r2_k1_train_full -f python scripts/prof_k1_traffic.py 256 train > gpurun_out/r2_k1_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
