"""stress K5-TC with train-loop-like calls (few frames, few / many ROIs, tiny boxes) to hunt a rare pipeline stall"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
torch.manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
H, W, C = 256, 448, 128
dev = "cuda"
feats4 = [torch.randn(4, C, H // s, W // s, device=dev) for s in (4, 8, 16, 32)]
t0 = time.time()
for it in range(iters):
    kind = it % 4
    N = 4
    if kind == 0:      # template boxes: 3 per frame, 30% of the image
        R = 12
        x1 = torch.rand(R, device=dev) * W * 0.6; y1 = torch.rand(R, device=dev) * H * 0.6
        rois = torch.stack([torch.arange(N, device=dev).repeat_interleave(3).float(), x1, y1, x1 + W * 0.3, y1 + H * 0.3], 1)
    elif kind == 1:    # 50 proposals per frame around a few objects
        R = 200
        x1 = torch.rand(R, device=dev) * W * 0.6; y1 = torch.rand(R, device=dev) * H * 0.6
        rois = torch.stack([torch.arange(N, device=dev).repeat_interleave(50).float(), x1, y1, x1 + W * 0.3 + 6 * torch.randn(R, device=dev), y1 + H * 0.3], 1)
    elif kind == 2:    # tiny boxes, some frames empty
        R = 7
        x1 = torch.rand(R, device=dev) * W; y1 = torch.rand(R, device=dev) * H
        rois = torch.stack([torch.randint(0, 2, (R,), device=dev).float() * 3, x1, y1, x1 + 5, y1 + 4], 1)
    else:              # random everything
        R = int(torch.randint(1, 150, (1,)))
        x1 = torch.rand(R, device=dev) * W; y1 = torch.rand(R, device=dev) * H
        rois = torch.stack([torch.randint(0, N, (R,), device=dev).float(), x1, y1, x1 + torch.rand(R, device=dev) * W, y1 + torch.rand(R, device=dev) * H], 1)
    out = ops.roi_mean_pool(feats4, rois, impl="tc")
    if it % 500 == 499:
        torch.cuda.synchronize()
        print("iter", it + 1, "ok", round(time.time() - t0, 1), "s", flush=True)
torch.cuda.synchronize()
print("done", iters)
