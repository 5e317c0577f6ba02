"""K3 solver timing at the bench shape (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
for B in (64, 888, 1024, 1036, 2048, 4096):
    sim = torch.rand(B, 10, 50, device="cuda")
    sc = torch.rand(B, 50, device="cuda")
    f = lambda: ops.relax_solve(sim, sc, None, None, 20, 5, 0.1, True, True, True)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    b.record(); torch.cuda.synchronize()
    print(f"K3 B={B}: {a.elapsed_time(b)/20*1e3:.1f} us")
