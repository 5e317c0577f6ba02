"""One K8 launch at N=3200 (for an ncu --set full capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
H, W, N = 256, 448, 3200
g = torch.Generator(device="cuda").manual_seed(0)
masks = torch.rand(N, 1, 28, 28, device="cuda", generator=g)
cx, cy = torch.rand(N, device="cuda", generator=g) * W, torch.rand(N, device="cuda", generator=g) * H
bw, bh = torch.rand(N, device="cuda", generator=g) * W * 0.5 + 8, torch.rand(N, device="cuda", generator=g) * H * 0.5 + 8
boxes = torch.stack([(cx - bw / 2).clamp(0, W - 1), (cy - bh / 2).clamp(0, H - 1), (cx + bw / 2).clamp(0, W - 1), (cy + bh / 2).clamp(0, H - 1)], 1)
for _ in range(2):
    ops.paste_masks(masks, boxes, H, W)
torch.cuda.synchronize()
