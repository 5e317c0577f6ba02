"""K5-TC pipeline trace of CTA 0 (debug build): prints event timeline in us relative to the first event"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
N, R = int(sys.argv[1]), int(sys.argv[2])
H, W, C = 256, 448, 128
g = torch.Generator(device="cuda").manual_seed(N)
feats = [torch.randn(N, C, H // s, W // s, generator=g, device="cuda") for s in (4, 8, 16, 32)]
x1 = torch.rand(N * R, generator=g, device="cuda") * W * 0.6
y1 = torch.rand(N * R, generator=g, device="cuda") * H * 0.6
bw = torch.rand(N * R, generator=g, device="cuda") * W * 0.3 + W / 8
bh = torch.rand(N * R, generator=g, device="cuda") * H * 0.3 + H / 8
rois = torch.stack([torch.arange(N, device="cuda").repeat_interleave(R).float(), x1, y1, (x1 + bw).clamp(max=W - 1), (y1 + bh).clamp(max=H - 1)], 1)
for _ in range(3): ops.roi_mean_pool(feats, rois, impl="tc")
torch.cuda.synchronize()
os.environ["DMM_K5_TRACE_FILE"] = "/tmp/k5trace.txt"
ops.roi_mean_pool(feats, rois, impl="tc")
torch.cuda.synchronize()
ev = [tuple(map(int, l.split())) for l in open("/tmp/k5trace.txt")]
ev.sort(key=lambda e: e[1])
t0 = ev[0][1]
names = {1: "P claim", 2: "P stream", 3: "P done", 4: "B start", 5: "B free", 6: "B built", 7: "M item", 8: "M bready", 9: "M last", 10: "E start", 11: "E rows", 12: "E fenced", 13: "E atom", 14: "E end"}
for e, t in ev[:400]:
    print(f"{(t - t0) / 1965.0:8.2f} us  {names.get(e, e)}")
