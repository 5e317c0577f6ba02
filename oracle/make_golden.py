#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Each file stores the seeded inputs and what the reference's own functions returned for them:
``dmm.modules.match_model.MatchModel`` (forward, compute_cost_matrix, match_with_first_frame),
``dmm.modules.submodules.relax_match.relax_matching`` and ``dmm.utils.match_helper``'s three helpers.
Train-mode cases also store autograd gradients through the reference layer.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from dmm.modules.match_model import MatchModel                      # noqa: E402  (the reference)
from dmm.modules.submodules.relax_match import relax_matching       # noqa: E402
from dmm.utils import match_helper as ref_helper                    # noqa: E402

from dmm_net_b200.synth import default_cfg, make_problem            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)                                            # one reduction order, reproducible


def npy(t):
    return t.detach().cpu().numpy()


def layer_case(name, P, O, H, W, D, max_iter, proj_iter, is_test, config, index, with_targets=False,
               dup_frac=0.0, mutate=None, lr=0.1, w=0.3):
    pr = make_problem(P, O, H, W, D, config=config, index=index, with_targets=with_targets, dup_frac=dup_frac)
    if mutate is not None:
        mutate(pr)
    cfg = default_cfg(max_iter, proj_iter, lr, w)
    layer = MatchModel(cfg, is_test=is_test)
    g = torch.Generator().manual_seed(77 + index)
    w_mask = torch.rand(O, H, W, generator=g)
    w_ms = torch.rand(O, generator=g)
    w_ds = torch.rand(O, generator=g)
    pf = pr.prop_feat.clone().requires_grad_(not is_test)
    tf = pr.tmpl_feat.clone().requires_grad_(not is_test)
    sc = pr.prop_score.clone().requires_grad_(not is_test)
    out = {}
    ctx = torch.no_grad() if is_test else torch.enable_grad()
    with ctx:
        sim, _, _, _ = layer.compute_cost_matrix({"proposed": pf, "template": [tf]},
                                                 {"proposed": pr.prop_mask, "template": pr.tmpl_mask},
                                                 {"proposal_score": sc}, pr.targets)
        _, _, _, logic, bmat = layer.match_with_first_frame(sim, P, O, pr.prop_mask.float(), sc, pr.tmpl_mask)
        sim_pad = sim.detach()
        if P <= O:
            sim_pad = torch.cat([sim_pad, sim_pad.new_zeros(O, O + 1 - P)], 1)
        _, _, X_list_ref, _ = relax_matching(-sim_pad, max_iter=max_iter, proj_iter=proj_iter, lr=lr)
        out["n_list"] = np.int64(len(X_list_ref))          # len(X_list): where the reference's outer exit fired
        full, ms, ds, full2, loss = layer(pf, pr.prop_mask, [tf], pr.tmpl_mask, sc, pr.targets)
        assert full is full2
        if not is_test:
            total = (full * w_mask).sum() + (ms * w_ms).sum() + (ds * w_ds).sum()
            if "cost_loss" in loss:
                total = total + 3.0 * loss["cost_loss"]
            total.backward()
            out.update(g_prop_feat=npy(pf.grad), g_tmpl_feat=npy(tf.grad),
                       g_prop_score=npy(sc.grad) if sc.grad is not None else np.zeros(P, np.float32))
    iou = ref_helper.compute_iou_binary_mask_2D(
        pr.prop_mask.view(P, -1).expand(O, -1, -1).contiguous().view(O * P, -1),
        pr.tmpl_mask.contiguous().view(O, 1, -1).expand(-1, P, -1).contiguous().view(O * P, -1)).view(O, P)
    cos = ref_helper.get_cosine_score(pr.tmpl_feat, pr.prop_feat)
    out.update(prop_feat=npy(pr.prop_feat), prop_mask=npy(pr.prop_mask), tmpl_feat=npy(pr.tmpl_feat),
               tmpl_mask=npy(pr.tmpl_mask), prop_score=npy(pr.prop_score),
               sim=npy(sim), iou=npy(iou), cos=npy(cos), logic=npy(logic), bmat=npy(bmat),
               full_outmask=npy(full), match_score=npy(ms), det_score=npy(ds),
               w_mask=npy(w_mask), w_ms=npy(w_ms), w_ds=npy(w_ds),
               meta=np.array([P, O, H, W, D, max_iter, proj_iter, is_test], np.int64),
               lr=np.float64(lr), score_weight=np.float64(w))
    if pr.targets is not None:
        out["targets"] = npy(pr.targets)
    if "cost_loss" in loss:
        out["cost_loss"] = npy(loss["cost_loss"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: sim[{O},{P}] full {tuple(full.shape)}  loss={loss.get('cost_loss', None)}")


def solver_case(name, C, max_iter, proj_iter, lr):
    C = torch.as_tensor(C, dtype=torch.float32)
    X, cost, X_list, inner = relax_matching(C, max_iter=max_iter, proj_iter=proj_iter, lr=lr)
    R = sum(X_list) / len(X_list)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), C=npy(C), X=npy(X), R=npy(R), X0=npy(X_list[0]),
                        n_list=np.int64(len(X_list)), cost=np.array(cost, np.float64),
                        params=np.array([max_iter, proj_iter], np.int64), lr=np.float64(lr))
    print(f"{name}: {tuple(C.shape)} len(X_list)={len(X_list)}")


def main():
    # ---- the layer ---------------------------------------------------------------------------------
    layer_case("layer_c1_test", 8, 3, 128, 128, 512, 20, 5, 1, config=1, index=0)           # BASELINE configs[0]
    layer_case("layer_c1_train", 8, 3, 64, 64, 512, 10, 5, 0, config=1, index=1, with_targets=True)
    layer_case("layer_eval40_odd", 13, 5, 37, 53, 64, 40, 5, 1, config=1, index=2)          # HW = 1961, odd
    layer_case("layer_pad_test", 3, 4, 40, 56, 32, 20, 5, 1, config=1, index=3)             # P <= O pad path
    layer_case("layer_pad_train", 4, 4, 24, 40, 32, 10, 5, 0, config=1, index=4, with_targets=True)
    layer_case("layer_dup_test", 10, 3, 32, 32, 16, 20, 5, 1, config=1, index=5, dup_frac=0.3)  # tied proposals

    def zero_some(pr):
        pr.tmpl_mask[0].zero_()          # union with an empty proposal is 0 -> IoU 0/1e-6 = 0
        pr.prop_mask[1].zero_()
        pr.prop_mask[2].fill_(0.5)       # exactly at the threshold: NOT set (strict >)
    layer_case("layer_zero_test", 6, 2, 16, 24, 8, 20, 5, 1, config=1, index=6, mutate=zero_some)
    layer_case("layer_c2_small", 50, 10, 64, 112, 512, 20, 5, 1, config=2, index=0)         # headline shape, small HW
    layer_case("layer_c2_train", 50, 5, 51, 64, 128, 10, 5, 0, config=2, index=1, with_targets=True)

    # ---- the solver alone --------------------------------------------------------------------------
    solver_case("solver_known_3x3", [[4, 1, 3], [2, 0, 5], [3, 2, 2]], 100, 100, 0.1)        # relax_match.py:108-119
    g = torch.Generator().manual_seed(5)
    solver_case("solver_10x50_20x5", -torch.rand(10, 50, generator=g), 20, 5, 0.1)
    solver_case("solver_10x50_40x5", -torch.rand(10, 50, generator=g), 40, 5, 0.1)
    solver_case("solver_5x6_10x5", -torch.rand(5, 6, generator=g), 10, 5, 0.1)
    solver_case("solver_1x2_20x5", -torch.rand(1, 2, generator=g), 20, 5, 0.1)
    solver_case("solver_16x64_default", -torch.rand(16, 64, generator=g), 400, 50, 0.1)     # configs/default.yaml
    solver_case("solver_zero_cost", torch.zeros(3, 7), 20, 5, 0.1)                           # outer exit at step 1
    solver_case("solver_greedy_only", -torch.rand(4, 9, generator=g), 0, 0, 0.0)             # match_helper.py:44
    solver_case("solver_7x33_longrun", -torch.rand(7, 33, generator=g), 300, 20, 0.05)

    # ---- helpers -----------------------------------------------------------------------------------
    a = torch.rand(9, 1000, generator=g)
    b = torch.rand(9, 1000, generator=g)
    a[3].zero_(); b[3].zero_()
    np.savez_compressed(os.path.join(OUT, "iou_rows.npz"), a=npy(a), b=npy(b),
                        iou=npy(ref_helper.compute_iou_binary_mask_2D(a, b)))
    q = torch.randn(7, 300, generator=g)
    k = torch.randn(21, 300, generator=g)
    k[4].zero_()                                                                              # zero vector -> eps clamp
    np.savez_compressed(os.path.join(OUT, "cosine.npz"), q=npy(q), k=npy(k), cos=npy(ref_helper.get_cosine_score(q, k)))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
