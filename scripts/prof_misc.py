"""Timings of the rows around the headline kernel: full layer (with apply), train-mode fwd+bwd, ROI mean pooling.
usage: python scripts/prof_misc.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
from dmm_net_b200.modules.match_model import MatchModel
from dmm_net_b200.synth import default_cfg, make_problems

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P, O, H, W, D = 50, 10, 256, 448, 512


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


pr = make_problems(B, P, O, H, W, D, seed=3, device="cuda", with_targets=True)
layer_t = MatchModel(default_cfg(20, 5), is_test=1)
with torch.no_grad():
    ms = timeit(lambda: layer_t.forward_many(pr.prop_feat, pr.prop_mask, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score))
full_bytes = B * ((P + O) * H * W * 4 + 2 * O * H * W * 4)
print(f"full layer (eval 20x5, +apply) B={B}: {ms:.3f} ms -> {B / ms * 1e3:.0f} matches/s, {full_bytes / ms / 1e6:.0f} GB/s (masks once + apply rows)")

one = pr.squeeze0()
with torch.no_grad():
    ms = timeit(lambda: layer_t(one.prop_feat, one.prop_mask, [one.tmpl_feat], one.tmpl_mask, one.prop_score), reps=20)
print(f"single-problem MatchModel.forward (the reference's per-video call, eval 20x5, incl. python + 5 launches): {ms * 1e3:.0f} us")

layer_tr = MatchModel(default_cfg(10, 5), is_test=0)
pf = pr.prop_feat.clone().requires_grad_(True)
tf = pr.tmpl_feat.clone().requires_grad_(True)


def train_step():
    out = layer_tr.forward_many(pf, pr.prop_mask, tf, pr.tmpl_mask, pr.prop_score, pr.targets)
    loss = out["full_outmask"].mean() + out["match_score"].sum() + out["cost_loss"].sum()
    loss.backward()
    pf.grad = None
    tf.grad = None


ms = timeit(train_step)
print(f"train-mode layer fwd+bwd (10x5, match loss, grads to features) B={B}: {ms:.3f} ms -> {B / ms * 1e3:.0f} matches/s")

# ROI mean pooling: 8 frames x 50 ROIs, 128 channels, 4 levels of a 256x448 image (config 3 shape)
N, C, R = 8, 128, 50
feats = [torch.randn(N, C, H // s, W // s, device="cuda") for s in (4, 8, 16, 32)]
g = torch.Generator(device="cuda").manual_seed(1)
x1 = torch.rand(N * R, device="cuda", generator=g) * (W - 40)
y1 = torch.rand(N * R, device="cuda", generator=g) * (H - 40)
w = 20 + torch.rand(N * R, device="cuda", generator=g) * (W / 2)
h = 20 + torch.rand(N * R, device="cuda", generator=g) * (H / 2)
rois = torch.stack([torch.arange(N, device="cuda").repeat_interleave(R).float(), x1, y1, (x1 + w).clamp(max=W - 1),
                    (y1 + h).clamp(max=H - 1)], 1)
with torch.no_grad():
    ms = timeit(lambda: ops.roi_mean_pool(feats, rois))
print(f"ROI mean pool fwd {N} frames x {R} ROIs x 4 levels x {C} ch: {ms * 1e3:.1f} us")
try:
    from torchvision.ops import roi_align
    def tv():
        outs = [roi_align(f, rois, (14, 14), spatial_scale=s, sampling_ratio=2, aligned=False).mean((2, 3))
                for f, s in zip(feats, (0.25, 0.125, 0.0625, 0.03125))]
        return torch.stack(outs, 1)
    with torch.no_grad():
        ms2 = timeit(tv)
    print(f"  torchvision roi_align x4 + mean on the same GPU (stand-in for maskrcnn_benchmark's ROIAlign): {ms2 * 1e3:.1f} us")
except Exception as e:
    print("  torchvision comparison skipped:", e)
fin = [f.clone().requires_grad_(True) for f in feats]
def pool_bwd():
    ops.roi_mean_pool(fin, rois).sum().backward()
    for f in fin:
        f.grad = None
ms = timeit(pool_bwd)
print(f"ROI mean pool fwd+bwd: {ms * 1e3:.1f} us")
