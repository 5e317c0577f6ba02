"""K4 assignment apply (eval mode: one selected proposal per template) -- CUDA events, GB/s of rows read + written."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
P, O, H, W = 50, 10, 256, 448
for B in (16, 64, 256, 512):
    prop = torch.rand(B, P, H * W, device="cuda")
    Bm = torch.zeros(B, O, 50, device="cuda")
    idx = torch.randint(0, P, (B, O), device="cuda")
    Bm.scatter_(2, idx[..., None], torch.rand(B, O, 1, device="cuda") + 0.1)
    f = lambda: ops.assign_apply(Bm, prop)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"K4 B={B}: {ms*1e3:.1f} us  {2*B*O*H*W*4/ms/1e6:.0f} GB/s", flush=True)
