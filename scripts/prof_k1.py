"""Small driver for ncu: a few launches of one kernel family on the headline shape.
usage: python scripts/prof_k1.py [k1|step|apply] [B] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
from dmm_net_b200.synth import make_problems

what = sys.argv[1] if len(sys.argv) > 1 else "k1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
P, O, H, W, D = 50, 10, 256, 448, 512
pr = make_problems(B, P, O, H, W, D, seed=1, device="cuda")
torch.cuda.synchronize()
with torch.no_grad():
    for i in range(reps):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        if what == "k1":
            r = ops.mask_iou_pairwise(pr.prop_mask, pr.tmpl_mask)
        elif what == "apply":
            cos = ops.cosine_pairwise(pr.tmpl_feat[:, None], pr.prop_feat)
            r = ops.mask_iou_pairwise(pr.prop_mask, pr.tmpl_mask, cos=cos, w_cos=0.7, w_iou=0.3)
            out = ops.relax_solve(r["sim"], pr.prop_score, max_iter=20, proj_iter=5, lr=0.1)
            t0.record()
            full = ops.assign_apply(out[1], pr.prop_mask, out[5])
        else:
            cos = ops.cosine_pairwise(pr.tmpl_feat[:, None], pr.prop_feat)
            r = ops.mask_iou_pairwise(pr.prop_mask, pr.tmpl_mask, cos=cos, w_cos=0.7, w_iou=0.3)
            out = ops.relax_solve(r["sim"], pr.prop_score, max_iter=20, proj_iter=5, lr=0.1)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        print(f"{what} B={B} rep{i}: {ms:.3f} ms  -> {B / ms * 1e3:.0f} matches/s, mask GB/s {B * 27525120 / ms / 1e6:.0f}")
