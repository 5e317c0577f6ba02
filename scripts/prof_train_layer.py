"""the matching layer in training mode (train.yaml 10x5, targets -> match loss), forward + backward at B problems:
CUDA-event timing per step, and the target of the ncu captures of the backward kernels (profiles/r2_train_bwd_ncu.txt)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200.modules.match_model import MatchModel
from dmm_net_b200.synth import default_cfg, make_problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pr = make_problems(B, 50, 10, 256, 448, 512, seed=3, device="cuda", with_targets=True)
layer = MatchModel(default_cfg(10, 5), is_test=0)
pf = pr.prop_feat.clone().requires_grad_(True)
tf = pr.tmpl_feat.clone().requires_grad_(True)
def step():
    out = layer.forward_many(pf, pr.prop_mask, tf, pr.tmpl_mask, pr.prop_score, pr.targets)
    (out["full_outmask"].mean() + out["match_score"].sum() + out["cost_loss"].sum()).backward()
    pf.grad = tf.grad = None
for _ in range(2): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): step()
b.record(); torch.cuda.synchronize()
print(f"train-mode layer fwd+bwd B={B}: {a.elapsed_time(b) / 3:.3f} ms per step, {B / (a.elapsed_time(b) / 3) * 1e3:.0f} matches/s")
