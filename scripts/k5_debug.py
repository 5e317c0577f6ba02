"""debug driver for K5-TC: one small case, blocking launches"""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
torch.manual_seed(0)
N, H, W, C = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 256, 448, 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 5
feats = [torch.randn(N, C, H // s, W // s, device="cuda") for s in (4, 8, 16, 32)]
x1 = torch.rand(N * R, device="cuda") * W * 0.6
y1 = torch.rand(N * R, device="cuda") * H * 0.6
rois = torch.stack([torch.arange(N, device="cuda").repeat_interleave(R).float(), x1, y1, x1 + 60, y1 + 50], 1)
a = ops.roi_mean_pool(feats, rois, impl="simt")
torch.cuda.synchronize(); print("simt ok", flush=True)
b = ops.roi_mean_pool(feats, rois, impl="tc")
torch.cuda.synchronize(); print("tc ok; max abs diff", float((a - b).abs().max()), flush=True)
for l in range(4):
    print("level", l, float((a[:, l * C:(l + 1) * C] - b[:, l * C:(l + 1) * C]).abs().max()))
