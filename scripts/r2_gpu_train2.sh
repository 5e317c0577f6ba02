#!/bin/bash
# round 2: train-step exchange variants on N GPUs (N = $1)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
for red in none flat ddp; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    examples/synthetic_train_step.py --steps 10 --warmup 3 --reduce $red 2>&1 | grep -v Warning | tail -2 | tee -a gpurun_out/r2_train_n${N}.txt
done
nproc; cat /sys/fs/cgroup/cpu.max
