"""one K1 launch at the headline shape for the ncu traffic capture (profiles/k1_traffic.json): B problems resident"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
from dmm_net_b200.synth import make_problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
train = len(sys.argv) > 2 and sys.argv[2] == "train"
pr = make_problems(B, 50, 10, 256, 448, 8, seed=1, device="cuda", with_targets=train)
for _ in range(3):
    r = ops.mask_iou_pairwise(pr.prop_mask, pr.tmpl_mask, pr.targets if train else None)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    r = ops.mask_iou_pairwise(pr.prop_mask, pr.tmpl_mask, pr.targets if train else None)
b.record(); torch.cuda.synchronize()
rows = 70 if train else 60
ms = a.elapsed_time(b) / 5
print(f"K1 {'train (70 rows, one pass)' if train else 'eval (60 rows)'} B={B}: {ms:.3f} ms  {rows * 256 * 448 * 4 * B / ms / 1e6:.0f} GB/s")
