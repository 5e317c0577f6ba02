"""K5 ROI mean pooling, forward: CUDA events at the per-frame and the batched sizes, plus a huge-ROI case (chunked table)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
H, W, C = 256, 448, 128
for N, R in ((8, 50), (64, 50)):
    g = torch.Generator(device="cuda").manual_seed(N)
    feats = [torch.randn(N, C, H // s, W // s, generator=g, device="cuda") for s in (4, 8, 16, 32)]
    x1 = torch.rand(N * R, generator=g, device="cuda") * W * 0.6
    y1 = torch.rand(N * R, generator=g, device="cuda") * H * 0.6
    bw = torch.rand(N * R, generator=g, device="cuda") * W * 0.3 + W / 8
    bh = torch.rand(N * R, generator=g, device="cuda") * H * 0.3 + H / 8
    rois = torch.stack([torch.arange(N, device="cuda").repeat_interleave(R).float(), x1, y1, (x1 + bw).clamp(max=W - 1), (y1 + bh).clamp(max=H - 1)], 1)
    big = rois.clone(); big[:, 1:] = torch.tensor([0.0, 0.0, W - 1.0, H - 1.0], device="cuda")
    for name, r in (("typical", rois), ("whole-image ROIs", big)):
        f = lambda: ops.roi_mean_pool(feats, r)
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): f()
        b.record(); torch.cuda.synchronize()
        print(f"K5 fwd {N} frames x {R} ROIs ({name}): {a.elapsed_time(b)/10*1e3:.1f} us", flush=True)
