#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for st in 5 4; do
  DMM_BUILD_DEFINES="-DK5_STAGES=$st -DK5_ALLOW_ODD_RING" python -m dmm_net_b200.build --force > /dev/null 2>&1
  echo "== raw ring of $st stages"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 scripts/k5_stress_nccl.py 6000 2>&1 | grep -E "iter|done|libdmm|Error" | sort | uniq -c | sort -rn | head -8
done
python -m dmm_net_b200.build --force > /dev/null 2>&1
