#!/usr/bin/env python
"""Golden vectors that need more of the reference than `make_golden.py` imports (build container only):

* `container_*.npz` -- the UNMODIFIED reference `DMM_Model` (dmm/modules/dmm_model.py:22-158: fill_template_dict,
  forward, inference, prepare_tplt_feature) run end to end.  The module imports `maskrcnn_benchmark` (un-vendored), so
  the ONE name it needs from there -- `modeling.poolers.Pooler` -- is stubbed with torchvision's legacy ROIAlign
  (`roi_align(aligned=False)`, the stand-in SURVEY.md section 8c names); everything else (the reference's own
  FeatureExtractor, the per-video loop, the 0/1-matrix select / scatter, MatchModel underneath) is the reference's code.
  Cases cover non-prefix `valid` rows, videos without templates, `extra_frame`, ragged proposal counts and training
  mode with gradients into the backbone features.
* `hun_*.npz` -- `MatchModel` with `algo='hun'` (relax_match.py:120-126).  `hungarian_matching` ends in `.cuda()`;
  the generator runs it on the CPU by making `Tensor.cuda` the identity for the duration of the call.
* `big_*.npz` -- the layer at the HEADLINE size (50 x 10 x 256x448, and 255x448).  The inputs (27.5 MB) are NOT stored:
  they come from the seeded generator (`dmm_net_b200.synth.make_problem`), and the file keeps their sha256 so the test
  can prove it regenerated the same bytes the reference saw; outputs are stored whole except `full_outmask`, which is
  stored on a pixel lattice (every 7th pixel) plus per-row sums.

    python oracle/make_golden_container.py
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)

from dmm_net_b200.synth import default_cfg, make_problem, make_problems      # noqa: E402
from dmm_net_b200.utils.boxlist import BoxList                                # noqa: E402  (duck-typed BoxList stand-in)


def npy(t):
    return t.detach().cpu().numpy()


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(npy(t)).tobytes())
    return h.hexdigest()


def import_ref_dmm_model():
    """reference dmm_model.py + feature_extractor.py with maskrcnn_benchmark's Pooler stubbed by torchvision ROIAlign."""
    from torchvision.ops import roi_align

    class _Level:
        def __init__(self, output_size, scale, sampling_ratio):
            self.output_size, self.scale, self.sampling_ratio = output_size, scale, sampling_ratio

        def __call__(self, feature, rois):
            return roi_align(feature, rois, self.output_size, spatial_scale=self.scale,
                             sampling_ratio=self.sampling_ratio, aligned=False)

    class Pooler:                                                   # maskrcnn_benchmark.modeling.poolers.Pooler
        def __init__(self, output_size, scales, sampling_ratio):
            self.poolers = [_Level(output_size, s, sampling_ratio) for s in scales]

    mb = types.ModuleType("maskrcnn_benchmark")
    modeling = types.ModuleType("maskrcnn_benchmark.modeling")
    poolers = types.ModuleType("maskrcnn_benchmark.modeling.poolers")
    poolers.Pooler = Pooler
    for n, m in (("maskrcnn_benchmark", mb), ("maskrcnn_benchmark.modeling", modeling),
                 ("maskrcnn_benchmark.modeling.poolers", poolers)):
        sys.modules.setdefault(n, m)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from dmm.modules import dmm_model
    return dmm_model


def _feats(gen, N, C, H, W):
    return [torch.randn(N, C, H // s, W // s, generator=gen) for s in (4, 8, 16, 32)]


def _boxes(gen, n, H, W):
    x1 = torch.rand(n, generator=gen) * (W - 8)
    y1 = torch.rand(n, generator=gen) * (H - 8)
    w = 4 + torch.rand(n, generator=gen) * (W / 2)
    h = 4 + torch.rand(n, generator=gen) * (H / 2)
    return torch.stack([x1, y1, (x1 + w).clamp(max=W - 1), (y1 + h).clamp(max=H - 1)], 1)


def container_case(name, ref_mod, is_test, n_prop, valid, extra_frame, seed, max_iter, H=64, W=96, C=16, Fm=5):
    B, P = len(n_prop), max(n_prop)
    gen = torch.Generator().manual_seed(seed)
    pr = make_problems(B, P, Fm, H, W, 4 * C, seed=seed + 1, with_targets=True)
    feats = [f.requires_grad_(not is_test) for f in _feats(gen, B, C, H, W)]
    tfeats = _feats(gen, B, C, H, W)                                # frame 0 (templates come from another frame)
    boxes = [_boxes(gen, n, H, W) for n in n_prop]
    tboxes = [_boxes(gen, Fm, H, W) for _ in range(B)]
    valid = torch.tensor(valid, dtype=torch.float32)
    cfg = default_cfg(max_iter, 5)
    model = ref_mod.DMM_Model(cfg, is_test=is_test)
    props = []
    for b in range(B):
        bl = BoxList(boxes[b], (W, H))
        bl.add_field("mask", pr.prop_mask[b, :n_prop[b]].unsqueeze(1))
        bl.add_field("scores", pr.prop_score[b, :n_prop[b]])
        props.append(bl)
    tplt = model.fill_template_dict(None, [BoxList(tb, (W, H)) for tb in tboxes],
                                    {"backbone_feature": tuple(tfeats), "refine_input_feat": tuple(tfeats)}, None, valid)
    tmpl_pooled = torch.stack([tplt[b]["feat"][0] for b in range(B)], 0)                 # [B,F,4C] (stand-in ROIAlign)
    out = {}
    if is_test:
        with torch.no_grad():
            om, _, loss, last = model.inference({"args": None, "shape": (H, W), "extra_frame": extra_frame, "valid": valid},
                                                props, tuple(feats), pr.tmpl_mask, tplt)
        assert loss == []
    else:
        om, _, loss, last = model(None, props, tuple(feats), pr.tmpl_mask, tplt, valid, pr.targets)
        g = torch.Generator().manual_seed(seed + 2)
        w_mask = torch.rand(om.shape, generator=g)
        total = (om * w_mask).sum() + 3.0 * sum(loss)
        total.backward()
        out.update(w_mask=npy(w_mask), match_loss=np.array([float(x) for x in loss], np.float32),
                   **{f"g_feat{l}": npy(feats[l].grad) for l in range(4)})
    pooled = model.feature_extractor(tuple(f.detach() for f in feats), props)
    out.update(output_mask=npy(om), out_mask_last=npy(last), valid=npy(valid), n_prop=np.array(n_prop, np.int64),
               extra_frame=np.array(extra_frame, np.int64), prop_mask=npy(pr.prop_mask), prop_score=npy(pr.prop_score),
               tmpl_mask=npy(pr.tmpl_mask), targets=npy(pr.targets), tmpl_pooled=npy(tmpl_pooled), prop_pooled=npy(pooled),
               meta=np.array([B, P, Fm, H, W, C, max_iter, 5, is_test], np.int64),
               **{f"feat{l}": npy(feats[l]) for l in range(4)}, **{f"tfeat{l}": npy(tfeats[l]) for l in range(4)},
               **{f"boxes{b}": npy(boxes[b]) for b in range(B)}, **{f"tboxes{b}": npy(tboxes[b]) for b in range(B)})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: output_mask {tuple(om.shape)} valid rows {valid.sum(1).tolist()} loss {[float(x) for x in loss]}")


def hun_case(name, P, O, H, W, D, is_test, index):
    from dmm.modules.match_model import MatchModel
    pr = make_problem(P, O, H, W, D, config=3, index=index)
    cfg = default_cfg(20, 5, algo="hun")
    layer = MatchModel(cfg, is_test=is_test)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self                  # relax_match.py:125 ends in .cuda(); CPU box here
    try:
        with torch.no_grad():
            sim, _, _, _ = layer.compute_cost_matrix({"proposed": pr.prop_feat, "template": [pr.tmpl_feat]},
                                                     {"proposed": pr.prop_mask, "template": pr.tmpl_mask},
                                                     {"proposal_score": pr.prop_score}, None)
            _, _, _, logic, bmat = layer.match_with_first_frame(sim, P, O, pr.prop_mask.float(), pr.prop_score, pr.tmpl_mask)
            full, ms, ds, _, _ = layer(pr.prop_feat, pr.prop_mask, [pr.tmpl_feat], pr.tmpl_mask, pr.prop_score)
    finally:
        torch.Tensor.cuda = orig
    np.savez_compressed(os.path.join(OUT, name + ".npz"), prop_feat=npy(pr.prop_feat), prop_mask=npy(pr.prop_mask),
                        tmpl_feat=npy(pr.tmpl_feat), tmpl_mask=npy(pr.tmpl_mask), prop_score=npy(pr.prop_score),
                        sim=npy(sim), logic=npy(logic), bmat=npy(bmat), full_outmask=npy(full), match_score=npy(ms),
                        det_score=npy(ds), meta=np.array([P, O, H, W, D, is_test], np.int64))
    print(f"{name}: hun assignment {bmat.argmax(1).tolist()}")


LATTICE = 7


def big_case(name, H, W, max_iter, index):
    from dmm.modules.match_model import MatchModel
    from dmm.modules.submodules.relax_match import relax_matching
    P, O, D = 50, 10, 512
    pr = make_problem(P, O, H, W, D, config=2, index=index)
    cfg = default_cfg(max_iter, 5)
    layer = MatchModel(cfg, is_test=1)
    torch.set_num_threads(8)
    with torch.no_grad():
        sim, _, _, _ = layer.compute_cost_matrix({"proposed": pr.prop_feat, "template": [pr.tmpl_feat]},
                                                 {"proposed": pr.prop_mask, "template": pr.tmpl_mask},
                                                 {"proposal_score": pr.prop_score}, None)
        _, _, X_list, _ = relax_matching(-sim, max_iter=max_iter, proj_iter=5, lr=0.1)
        _, _, _, logic, bmat = layer.match_with_first_frame(sim, P, O, pr.prop_mask.float(), pr.prop_score, pr.tmpl_mask)
        full, ms, ds, _, _ = layer(pr.prop_feat, pr.prop_mask, [pr.tmpl_feat], pr.tmpl_mask, pr.prop_score)
    torch.set_num_threads(1)
    flat = full.reshape(O, -1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sim=npy(sim), logic=npy(logic), bmat=npy(bmat),
                        match_score=npy(ms), det_score=npy(ds), n_list=np.int64(len(X_list)),
                        full_lattice=npy(flat[:, ::LATTICE]), full_rowsum=npy(flat.double().sum(1)),
                        input_sha=np.array(sha(pr.prop_feat, pr.prop_mask, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score)),
                        meta=np.array([P, O, H, W, D, max_iter, 5, 1, 2, index, LATTICE], np.int64))
    print(f"{name}: {H}x{W} len(X_list)={len(X_list)} assignment {bmat.argmax(1).tolist()}")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_mod = import_ref_dmm_model()
    # inference: non-prefix valid rows (video 0), one template (1), none (2), all (3); video 1 is an `extra_frame`
    container_case("container_eval", ref_mod, 1, [12, 7, 12, 3], [[1, 0, 1, 1, 0], [1, 0, 0, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]],
                   [0, 0, 0, 0], seed=301, max_iter=40)
    container_case("container_eval_extra", ref_mod, 1, [9, 12, 5], [[1, 1, 0, 0, 0], [1, 1, 1, 0, 0], [0, 1, 1, 0, 1]],
                   [0, 1, 0], seed=302, max_iter=20)
    # training: forward() with targets, gradients into the four backbone feature levels
    container_case("container_train", ref_mod, 0, [12, 6, 12], [[1, 1, 1, 0, 0], [0, 0, 0, 0, 0], [1, 1, 0, 0, 0]],
                   [0, 0, 0], seed=303, max_iter=10)
    hun_case("hun_test", 8, 3, 64, 64, 64, 1, 0)
    hun_case("hun_pad_test", 3, 4, 40, 56, 32, 1, 1)               # P <= O: zero-padded to O+1 columns
    hun_case("hun_train", 9, 4, 32, 48, 32, 0, 2)
    big_case("big_c2_full", 256, 448, 20, 0)                       # BASELINE configs[1], exactly
    big_case("big_c2_odd", 255, 448, 40, 1)                        # the real scripts' 255x448 with eval.yaml's 40x5
    print("written to", OUT)


if __name__ == "__main__":
    main()
