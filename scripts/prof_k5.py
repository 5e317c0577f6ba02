"""K5 ROI mean pooling: CUDA-event timing of the tensor-core path (roi_pool_tc.cu) and the gather kernel, forward and
backward, at the per-clip (8 frames) and batched (64 frames) sizes; algorithmic bytes = every feature byte once + out."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
H, W, C = 256, 448, 128
impls = sys.argv[1:] or ["tc", "simt"]
for N, R in ((8, 50), (64, 50)):
    g = torch.Generator(device="cuda").manual_seed(N)
    feats = [torch.randn(N, C, H // s, W // s, generator=g, device="cuda") for s in (4, 8, 16, 32)]
    x1 = torch.rand(N * R, generator=g, device="cuda") * W * 0.6
    y1 = torch.rand(N * R, generator=g, device="cuda") * H * 0.6
    bw = torch.rand(N * R, generator=g, device="cuda") * W * 0.3 + W / 8
    bh = torch.rand(N * R, generator=g, device="cuda") * H * 0.3 + H / 8
    rois = torch.stack([torch.arange(N, device="cuda").repeat_interleave(R).float(), x1, y1, (x1 + bw).clamp(max=W - 1), (y1 + bh).clamp(max=H - 1)], 1)
    big = rois.clone(); big[:, 1:] = torch.tensor([0.0, 0.0, W - 1.0, H - 1.0], device="cuda")
    few = rois[::10].contiguous()                                    # 5 ROIs per frame (template pooling)
    algo = sum(f.numel() for f in feats) * 4 + N * R * 4 * C * 4
    for name, r in (("typical", rois), ("whole-image ROIs", big), ("5 ROIs per frame", few)):
        for impl in impls:
            f = lambda: ops.roi_mean_pool(feats, r, impl=impl)
            for _ in range(3): f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): f()
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 10 * 1e3
            print(f"K5 fwd [{impl:4s}] {N} frames x {r.shape[0] // N} ROIs ({name}): {us:.1f} us  ({algo / us / 1e6:.2f} TB/s of feature bytes once)", flush=True)
    if "--bwd" in sys.argv or True:
        fin = [f.clone().requires_grad_(True) for f in feats]
        out = ops.roi_mean_pool(fin, rois, impl=impls[0])
        go = torch.randn_like(out)
        for _ in range(2): torch.autograd.grad(out, fin, go, retain_graph=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): torch.autograd.grad(out, fin, go, retain_graph=True)
        b.record(); torch.cuda.synchronize()
        print(f"K5 bwd {N} frames x {R} ROIs: {a.elapsed_time(b) / 5 * 1e3:.1f} us", flush=True)
