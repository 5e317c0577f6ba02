#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DMM_BUILD_DEFINES="-DDMM_TC_DEBUG" python -m dmm_net_b200.build --force > /dev/null 2>&1
echo "== K5 on cuda:1 (single process)"
timeout 200 python - <<'PY' 2>&1 | grep -v "^  File\|^    \|Warning" | sort | uniq -c | sort -rn | head -12
import torch, sys
sys.path.insert(0, ".")
torch.cuda.set_device(1)
from dmm_net_b200 import ops
N, R, H, W, C = 4, 50, 256, 448, 128
g = torch.Generator(device="cuda:1").manual_seed(1)
feats = [torch.randn(N, C, H // s, W // s, generator=g, device="cuda:1") for s in (4, 8, 16, 32)]
x1 = torch.rand(N * R, generator=g, device="cuda:1") * W * 0.6
y1 = torch.rand(N * R, generator=g, device="cuda:1") * H * 0.6
rois = torch.stack([torch.arange(N, device="cuda:1").repeat_interleave(R).float(), x1, y1, x1 + 80, y1 + 60], 1)
for i in range(20):
    a = ops.roi_mean_pool(feats, rois, impl="tc")
torch.cuda.synchronize()
b = ops.roi_mean_pool(feats, rois, impl="simt")
print("cuda:1 tc ok, max diff", float((a - b).abs().max()))
PY
echo "== 2-rank train, flat"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 examples/synthetic_train_step.py --steps 30 --warmup 2 --reduce flat 2>&1 | grep -v "^  File\|^    \|Warning\|^\[rank" | sort | uniq -c | sort -rn | head -14
python -m dmm_net_b200.build --force > /dev/null 2>&1
