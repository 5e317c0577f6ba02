"""One launch each of K4 / K6 / K7 / K8 / K10 at batch sizes where they stream, for an ncu capture (profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
H, W, P, O, B = 256, 448, 50, 10, 128
prop = torch.rand(B, P, H * W, device="cuda")
Bm = torch.zeros(B, O, 50, device="cuda")
Bm.scatter_(2, torch.randint(0, P, (B, O, 1), device="cuda"), 1.0)
for _ in range(2):
    ops.assign_apply(Bm, prop)                                              # K4
    a, b, c = (torch.rand(B, O, H, W, device="cuda") for _ in range(3))
    ops.mask_pyramid(a, b, c, 4)                                            # K6
    ops.merge_labels(a.view(B, O, -1))                                      # K7
    m28 = torch.rand(B * 25, 1, 28, 28, device="cuda")
    xy = torch.rand(B * 25, 2, device="cuda") * torch.tensor([W * 0.6, H * 0.6], device="cuda")
    bx = torch.cat([xy, xy + torch.tensor([W * 0.3, H * 0.3], device="cuda")], 1)
    ops.paste_masks(m28, bx, H, W, want_bits=True)                          # K8
    src = torch.randint(0, B * 25, (B, P), device="cuda", dtype=torch.int32)
    ops.paste_apply(Bm, m28, bx, src, H, W)                                 # K10
torch.cuda.synchronize()
print("ok")
