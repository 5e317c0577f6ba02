"""CPU oracle for the DMM-Net matching hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch-CPU fp32 tensors, no CUDA) of the
reference algorithm.  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product path
(``dmm_net_b200``) never imports it and fails loudly without its CUDA library.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (it runs in the build container), feeds both with identical seeded
inputs and stores the reference outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
(bit-exact for IoU / greedy init, <=1e-6 for the float paths) and against the reference's
single known answer (``relax_match.py:108-119``: the 3x3 cost whose relaxed solution equals
SciPy's Hungarian assignment).

Why torch-CPU and not numpy: the reference *is* torch fp32; using the same primitive ops
in the same order makes this port bit-identical to the reference on the same host, and makes
its wall-clock a faithful stand-in for "the reference's CPU path" on the GPU box where
``/root/reference`` does not exist.

Reference files restated (all under /root/reference/dmm):
  utils/match_helper.py:9-28    -> rowwise_binary_iou
  utils/match_helper.py:51-64   -> cosine_scores
  utils/match_helper.py:30-49   -> matching_loss
  modules/submodules/relax_match.py:9-19,21-34 -> _row_projection, _col_projection
  modules/submodules/relax_match.py:45-55      -> greedy_init
  modules/submodules/relax_match.py:36-105     -> relax_solve
  modules/submodules/relax_match.py:120-126    -> hungarian_onehot
  modules/match_model.py:49-91   -> cost_matrix
  modules/match_model.py:93-148  -> assign_and_apply
  modules/match_model.py:24-47   -> match_layer_forward
  modules/feature_extractor.py:20-52 (+ maskrcnn_benchmark ROIAlign, un-vendored) -> roi_mean_pool
  modules/dmm_model.py:88-158    -> dmm_container_forward
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "rowwise_binary_iou", "pairwise_binary_iou", "cosine_scores", "greedy_init",
    "relax_solve", "hungarian_onehot", "matching_loss", "cost_matrix",
    "assign_and_apply", "match_layer_forward", "roi_mean_pool", "roi_mean_pool_separable",
    "dmm_container_forward",
]


# --------------------------------------------------------------------------------------
# mask IoU  (match_helper.py:9-28)
# --------------------------------------------------------------------------------------
def rowwise_binary_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Row i of ``a`` against row i of ``b``: threshold at 0.5, |A&B| / (|A|B| + 1e-6).

    The two sums are fp32 sums of 0/1 values (exact integers below 2**24), the epsilon is
    added in fp32 and the divide is an IEEE fp32 divide -- so the result is reproducible
    bit for bit by any implementation that counts exactly (match_helper.py:16-27).
    """
    assert a.dim() == 2 and b.dim() == 2, (a.shape, b.shape)
    bits_a = (a > 0.5).to(torch.uint8)
    bits_b = (b > 0.5).to(torch.uint8)
    with torch.no_grad():
        union = (bits_a | bits_b).float().sum(1) + 1e-6
        inter = (bits_a & bits_b).float().sum(1)
        out = inter / union
    return out.detach()


def pairwise_binary_iou(prop: torch.Tensor, tmpl: torch.Tensor, expand: bool = True) -> torch.Tensor:
    """[P,HW] proposals x [O,HW] templates -> [O,P] IoU (match_model.py:83-89).

    ``expand=True`` reproduces the reference's memory behaviour (two materialised
    [O*P, HW] copies) -- that is what the CPU baseline must time.  ``expand=False``
    computes the same integers from three popcount-style sums (used for big test cases).
    """
    P, O = prop.shape[0], tmpl.shape[0]
    if expand:
        pe = prop.reshape(P, -1).expand(O, -1, -1).contiguous().view(O * P, -1)
        te = tmpl.contiguous().view(O, 1, -1).expand(-1, P, -1).contiguous().view(O * P, -1)
        return rowwise_binary_iou(pe, te).view(O, P)
    pb = (prop.reshape(P, -1) > 0.5).float()
    tb = (tmpl.reshape(O, -1) > 0.5).float()
    inter = tb @ pb.t()                                   # exact: 0/1 products, counts < 2**24
    union = tb.sum(1, keepdim=True) + pb.sum(1)[None, :] - inter
    return inter / (union + 1e-6)


# --------------------------------------------------------------------------------------
# cosine  (match_helper.py:51-64)
# --------------------------------------------------------------------------------------
def cosine_scores(query: torch.Tensor, key: torch.Tensor) -> torch.Tensor:
    """[O,D] x [P,D] -> [O,P] cosine similarity, eps = 1e-8 (F.cosine_similarity default)."""
    O, P = query.shape[0], key.shape[0]
    q = query.unsqueeze(2).expand(-1, -1, P)              # O,D,P
    k = key.permute(1, 0).expand(O, -1, -1)               # O,D,P
    assert q.shape == k.shape
    return F.cosine_similarity(q, k, dim=1)


# --------------------------------------------------------------------------------------
# relaxed matching  (relax_match.py)
# --------------------------------------------------------------------------------------
def greedy_init(C: torch.Tensor) -> torch.Tensor:
    """One-hot start of the solver (relax_match.py:45-55), vectorised.

    Every column keeps only the entry of its best (lowest-cost, first on ties) row, the
    rest is overwritten with max(C); then every row picks its lowest column (first on ties).
    """
    n, m = C.shape
    fill = C.max()
    best_row = torch.argmin(C, dim=0)                     # per column
    kept = torch.full_like(C, fill.item())
    cols = torch.arange(m)
    kept[best_row, cols] = C[best_row, cols]
    pick = torch.min(kept, dim=1)[1]
    X = torch.zeros_like(C)
    X[torch.arange(n), pick] = 1.0
    return X


def _row_projection(X: torch.Tensor) -> torch.Tensor:
    """{row sums == 1}: X - (rowsum - 1)/m   (relax_match.py:9-19)."""
    return X - (X.sum(dim=1, keepdim=True) - 1) / X.shape[1]


def _col_projection(X: torch.Tensor) -> torch.Tensor:
    """{col sums <= 1}: only columns whose sum exceeds 1 move (relax_match.py:21-34)."""
    s = X.sum(dim=0, keepdim=True)
    keep = (s <= 1).float()
    moved = X - (s - 1).expand_as(X) / X.shape[0]
    return X * keep + (1 - keep) * moved


def relax_solve(C: torch.Tensor, max_iter: int = 100, proj_iter: int = 100, lr: float = 0.1,
                force_len: Optional[int] = None):
    """Projected gradient descent with Dykstra sweeps (relax_match.py:36-105).

    ``force_len`` (tests only): ignore the outer exact-equality exit and stop once ``len(X_list) == force_len``.
    The outer exit compares two fp32 norms for equality; once the iteration has converged the step at which that
    happens depends on summation order down to the last bit (the reference's own CPU/AVX2/AVX512/CUDA builds differ),
    so parity tests compare an implementation that stopped after L iterates with the reference algorithm stopped
    after the same L.

    Returns ``(X, cost, X_list, last_inner_errors)`` exactly like the reference:
    ``X_list[0]`` is the greedy start, ``X_list[k]`` the iterate right after gradient step k
    (i.e. BEFORE it is projected); the three Dykstra increments live across outer steps;
    the inner loop stops when a sweep changes nothing, the outer loop when two consecutive
    ``||X*C||`` agree exactly.
    """
    X = greedy_init(C)
    X_list = [X]
    inc = [torch.zeros_like(C) for _ in range(3)]
    cost = [0]
    inner_err: list = 0
    for _ in range(max_iter):
        if force_len is not None and len(X_list) >= force_len:
            break
        X = X - lr * C
        cost.append((X * C).norm().item())
        X_list.append(X)
        inner_err = []
        for _ in range(proj_iter):
            X0 = X.clone()
            X = X + inc[0]
            Y = F.relu(X)
            inc[0] = X - Y
            X = Y + inc[1]
            Y = _col_projection(X)
            inc[1] = X - Y
            X = Y + inc[2]
            Y = _row_projection(X)
            inc[2] = X - Y
            X = Y
            delta = (X - X0).norm().item()
            if delta == 0:
                break
            inner_err.append(delta)
        if force_len is None and cost[-2] == cost[-1]:
            break
    return X, cost, X_list, inner_err


def hungarian_onehot(C: torch.Tensor) -> torch.Tensor:
    """SciPy LSAP one-hot (relax_match.py:120-126), kept on the CPU."""
    from scipy.optimize import linear_sum_assignment
    r, c = linear_sum_assignment(C.detach().cpu().numpy())
    X = torch.zeros_like(C)
    X[torch.as_tensor(r), torch.as_tensor(c)] = 1.0
    return X


# --------------------------------------------------------------------------------------
# match loss  (match_helper.py:30-49)
# --------------------------------------------------------------------------------------
def matching_loss(prop_mask: torch.Tensor, targets: torch.Tensor, feature_sim: torch.Tensor,
                  expand: bool = True) -> torch.Tensor:
    assert prop_mask.dim() == 3 and targets.dim() == 3
    assert prop_mask.shape[-1] == targets.shape[-1]
    P, O = prop_mask.shape[0], targets.shape[0]
    gt_iou = pairwise_binary_iou((prop_mask > 0.5).float().view(P, -1), targets.reshape(O, -1), expand)
    gt_match = greedy_init(-gt_iou)                       # relax_matching(-iou, 0, 0, 0)[0]
    return F.mse_loss(feature_sim, gt_match)


# --------------------------------------------------------------------------------------
# the layer  (match_model.py)
# --------------------------------------------------------------------------------------
def cost_matrix(prop_feat, prop_mask, tmpl_feats: Sequence[torch.Tensor], tmpl_mask, score_weight: float,
                targets=None, expand: bool = True):
    """sim = (1-w)*mean_t cos + w*IoU ; optional match loss on the cosine part (match_model.py:49-91)."""
    assert prop_mask.dim() == 3
    P, O = prop_mask.shape[0], tmpl_feats[0].shape[0]
    fsim = prop_feat.new_zeros(O, P)
    for tf in tmpl_feats:
        fsim = fsim + cosine_scores(tf, prop_feat)
    fsim = fsim / len(tmpl_feats)
    loss: Dict[str, torch.Tensor] = {}
    if targets is not None:
        loss["cost_loss"] = matching_loss(prop_mask, targets, fsim, expand)
    iou = pairwise_binary_iou(prop_mask.reshape(P, -1), tmpl_mask.reshape(O, -1), expand)
    sim = fsim * (1 - score_weight) + iou * score_weight
    return sim, loss


def assign_and_apply(sim, prop_mask, prop_score, max_iter, proj_iter, lr, is_test: int, algo: str = "relax",
                     force_len: Optional[int] = None):
    """match_model.py:93-148.  Returns (full_outmask, match_score, det_score, logic, Bmat, R)."""
    O, P = sim.shape
    pad = 0
    if P <= O:
        sim_p = sim.new_zeros((O, O + 1))
        sim_p[:, :P] = sim
        pad = O + 1 - P
    else:
        sim_p = sim
    C = -sim_p
    if algo == "relax":
        _, _, X_list, _ = relax_solve(C, max_iter, proj_iter, lr, force_len)
        R = sum(X_list) / len(X_list)
    else:
        R = hungarian_onehot(C)
    top = R.max(dim=1, keepdim=True)[0]
    logic = (R == top).float() if is_test else (R > 0.01).float()
    Bm = R.float() * logic
    H, W = prop_mask.shape[-2:]
    m2 = prop_mask.float().reshape(P, -1)
    sc = prop_score
    if pad:
        m2 = torch.cat([m2, m2.new_zeros(pad, m2.shape[1])], 0)
        sc = torch.cat([prop_score, prop_score.new_zeros(pad)], 0)
    full = torch.mm(Bm, m2).view(-1, H, W)
    match_score = (R.clamp(0, 1) * (-C)).max(1)[0]
    det_score = (sc.view(1, -1).expand(O, -1) * Bm).sum(1)
    return full, match_score, det_score, logic, Bm, R


def match_layer_forward(cfg: dict, is_test: int, prop_feat, prop_mask, tmpl_feats, tmpl_mask, prop_score,
                        targets=None, expand: bool = True, force_len: Optional[int] = None):
    """MatchModel.forward (match_model.py:24-47).  cfg carries the five keys the reference reads."""
    algo = cfg["matching"]["algo"]
    assert algo in ("relax", "hun")
    sim, loss = cost_matrix(prop_feat, prop_mask, tmpl_feats, tmpl_mask, cfg["score_weight"], targets, expand)
    full, ms, ds, _, _, _ = assign_and_apply(sim, prop_mask.float(), prop_score, cfg["relax_max_iter"],
                                              cfg["relax_proj_iter"], cfg["relax_learning_rate"], is_test, algo, force_len)
    return full, ms, ds, full, loss


# --------------------------------------------------------------------------------------
# proposal-feature pooling  (feature_extractor.py:11-52 over maskrcnn_benchmark's ROIAlign)
# --------------------------------------------------------------------------------------
POOL_SCALES = (0.25, 0.125, 0.0625, 0.03125)      # feature_extractor.py:13
POOL_RES = 14                                     # feature_extractor.py:15
POOL_SAMPLING = 2                                 # feature_extractor.py:14


def roi_mean_pool(features: Sequence[torch.Tensor], rois: torch.Tensor) -> torch.Tensor:
    """Stand-in oracle (parity UNPINNED by the reference: the arithmetic lives in the
    un-vendored ZENGXH/maskrcnn-benchmark fork, no pinned commit).  Legacy ROIAlign ==
    torchvision.ops.roi_align(aligned=False); then mean over the 14x14 bins, levels concatenated.
    ``rois`` = [R,5] rows (batch_idx, x1, y1, x2, y2) (feature_extractor.py:32-37)."""
    from torchvision.ops import roi_align
    outs = []
    for f, s in zip(features, POOL_SCALES):
        pooled = roi_align(f, rois, (POOL_RES, POOL_RES), spatial_scale=s, sampling_ratio=POOL_SAMPLING, aligned=False)
        outs.append(pooled.mean(3).mean(2))                # == .mean(4).mean(3) on [R,L,C,h,w]
    return torch.stack(outs, 1).reshape(rois.shape[0], -1)


def _axis_weights(lo: float, hi: float, size: int, scale: float) -> torch.Tensor:
    """1-D weight vector w[size] such that mean over 14 bins x 2 samples of bilinear taps == w . f."""
    w = torch.zeros(size, dtype=torch.float64)
    start = lo * scale
    length = max(hi * scale - start, 1.0)
    bin_sz = length / POOL_RES
    for b in range(POOL_RES):
        for s in range(POOL_SAMPLING):
            t = start + b * bin_sz + (s + 0.5) * bin_sz / POOL_SAMPLING
            if t < -1.0 or t > size:
                continue
            t = max(t, 0.0)
            lo_i = int(t)
            if lo_i >= size - 1:
                lo_i = hi_i = size - 1
                t = float(lo_i)
            else:
                hi_i = lo_i + 1
            frac = t - lo_i
            w[lo_i] += 1.0 - frac
            w[hi_i] += frac
    return w / (POOL_RES * POOL_SAMPLING)


def roi_mean_pool_separable(features: Sequence[torch.Tensor], rois: torch.Tensor) -> torch.Tensor:
    """Same quantity as ``roi_mean_pool`` written as the separable contraction the CUDA kernel
    uses: out[r,l,c] = sum_y sum_x wy[y] wx[x] F_l[b,c,y,x]."""
    R = rois.shape[0]
    outs = []
    for f, s in zip(features, POOL_SCALES):
        _, Cc, Hh, Ww = f.shape
        o = torch.zeros(R, Cc, dtype=torch.float64)
        for r in range(R):
            b, x1, y1, x2, y2 = [float(v) for v in rois[r]]
            wy = _axis_weights(y1, y2, Hh, s)
            wx = _axis_weights(x1, x2, Ww, s)
            o[r] = torch.einsum("y,cyx,x->c", wy, f[int(b)].double(), wx)
        outs.append(o.float())
    return torch.stack(outs, 1).reshape(R, -1)


# --------------------------------------------------------------------------------------
# container  (dmm_model.py:88-158) -- the caller of the boundary, used by the "next" row tests
# --------------------------------------------------------------------------------------
def dmm_container_forward(cfg, is_test, prop_feats: List[torch.Tensor], prop_masks: List[torch.Tensor],
                          prop_scores: List[torch.Tensor], tmpl_feat: List[torch.Tensor], mask_last: torch.Tensor,
                          valid: torch.Tensor, targets: Optional[torch.Tensor] = None, expand: bool = False,
                          extra_frame: Optional[Sequence[int]] = None):
    """Per-video loop of DMM_Model.forward / .inference.

    prop_*[b]: tensors of video b; tmpl_feat[b]: [F,D]; mask_last: [B,F,H,W]; valid: [B,F] 0/1 with
    O = valid.sum() templates: the reference takes rows :O of diag(valid) and of the masks (dmm_model.py:125,152-154),
    whatever the positions of the ones.  ``extra_frame[b]`` (inference only, dmm_model.py:66) skips video b.
    Pinned to the reference by tests/golden/container_*.npz (oracle/make_golden_container.py).
    Returns (output_mask [B,F,H,W], match_loss list, out_mask_last [B,F,H,W])."""
    B, Fm, H, W = mask_last.shape
    outs, lasts, losses = [], [], []
    for b in range(B):
        v = valid[b]
        O = int(v.sum().item())
        if O == 0 or (extra_frame is not None and extra_frame[b]):
            outs.append(mask_last.new_zeros(Fm, H, W))
            lasts.append(mask_last[b])
            if not is_test:
                losses.append(prop_feats[b].sum() * 0)
            continue
        sel = torch.diag(v).float()[:O, :]                 # [O,F]
        tf = torch.mm(sel, tmpl_feat[b].view(Fm, -1))
        tg = None if targets is None else targets[b, :O].view(O, H, W)
        full, _, _, last, loss = match_layer_forward(cfg, is_test, prop_feats[b], prop_masks[b], [tf],
                                                     mask_last[b, :O].view(O, H, W), prop_scores[b], tg, expand)
        outs.append(torch.mm(sel.t(), full.view(O, -1)).view(Fm, H, W))
        lasts.append(torch.mm(sel.t(), last.view(O, -1)).view(Fm, H, W))
        if len(loss) > 0:
            losses.append(loss["cost_loss"])
    return torch.stack(outs, 0), losses, torch.stack(lasts, 0)
