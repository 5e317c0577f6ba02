"""GPU parity of the rows either side of the layer: K5 proposal-feature pooling (vs torchvision's legacy ROIAlign,
the stand-in oracle -- parity unpinned by the reference) and the batched DMM_Model container (vs the per-video loop
of the oracle's dmm_container_forward, reference dmm_model.py:88-158)."""
import numpy as np
import pytest
import torch

from dmm_net_b200 import ops
from dmm_net_b200.modules.dmm_model import DMM_Model
from dmm_net_b200.modules.feature_extractor import make_roi_mask_feature_extractor
from dmm_net_b200.synth import default_cfg, make_problems
from dmm_net_b200.utils.boxlist import BoxList
from oracle import match_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _feats(gen, N, C, H, W):
    return [torch.randn(N, C, H // s, W // s, generator=gen) for s in (4, 8, 16, 32)]


def _boxes(gen, n, H, W):
    x1 = torch.rand(n, generator=gen) * (W - 8)
    y1 = torch.rand(n, generator=gen) * (H - 8)
    w = 4 + torch.rand(n, generator=gen) * (W / 2)
    h = 4 + torch.rand(n, generator=gen) * (H / 2)
    return torch.stack([x1, y1, (x1 + w).clamp(max=W - 1), (y1 + h).clamp(max=H - 1)], 1)


def test_roi_mean_pool_forward_and_backward():
    gen = torch.Generator().manual_seed(4)
    N, C, H, W = 2, 24, 256, 448
    feats = _feats(gen, N, C, H, W)
    rois = torch.cat([torch.cat([torch.full((9, 1), float(i)), _boxes(gen, 9, H, W)], 1) for i in range(N)], 0)
    rois = torch.cat([rois, torch.tensor([[0, -30.0, -10.0, 20.0, 40.0], [1, 440.0, 250.0, 470.0, 280.0],
                                          [1, 100.0, 100.0, 100.0, 100.0], [0, 0.0, 0.0, 447.0, 255.0]])], 0)
    want_in = [f.clone().requires_grad_(True) for f in feats]
    want = orc.roi_mean_pool(want_in, rois)
    got_in = [f.to(DEV).requires_grad_(True) for f in feats]
    got = ops.roi_mean_pool(got_in, rois.to(DEV))
    assert got.shape == (rois.shape[0], 4 * C)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=0, atol=1e-5)
    w = torch.randn(want.shape, generator=gen)
    # reference gradient in float64: the stand-in's own fp32 backward adds up to 784 equal terms sequentially for the point
    # ROI (2e-5 off); the gather kernel multiplies the two 1-D weights instead
    want64_in = [f.double().requires_grad_(True) for f in feats]
    (orc.roi_mean_pool(want64_in, rois.double()) * w.double()).sum().backward()
    (got * w.to(DEV)).sum().backward()
    for a, b in zip(got_in, want64_in):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.float().numpy(), rtol=1e-5, atol=2e-6)   # deterministic gather


@pytest.mark.parametrize("N,C,H,W,counts", [(3, 128, 256, 448, [50, 0, 9]), (2, 40, 128, 192, [130, 3]), (2, 136, 64, 96, [5, 70])])
def test_roi_mean_pool_backward_is_a_deterministic_gather(N, C, H, W, counts):
    """K5 backward: every gradient element written once in a fixed order -- bit-identical run to run, equal (to fp32
    summation order) to the atomic scatter and to autograd through the ROIAlign stand-in; frames with several ROI
    groups (> 64 ROIs), a frame without ROIs (gradient exactly zero), any channel count."""
    gen = torch.Generator().manual_seed(7 * N + C)
    feats, rois = _tc_case(gen, N, H, W, counts, C=C)
    rois = rois[(rois[:, 0] >= 0) & (rois[:, 0] < N)]
    w = torch.randn(rois.shape[0], 4 * C, generator=gen)

    def grads(mode):
        ops.ROI_POOL_BWD_IMPL = mode
        try:
            fin = [f.to(DEV).requires_grad_(True) for f in feats]
            out = ops.roi_mean_pool(fin, rois.to(DEV))
            (out * w.to(DEV)).sum().backward()
            return [f.grad for f in fin]
        finally:
            ops.ROI_POOL_BWD_IMPL = "auto"

    a, b, c = grads("auto"), grads("auto"), grads("atomic")
    want_in = [f.double().requires_grad_(True) for f in feats]                        # float64 reference (see the test above)
    (orc.roi_mean_pool(want_in, rois.double()) * w.double()).sum().backward()
    for l in range(4):
        assert torch.equal(a[l], b[l]), l                                             # run to run
        scale = max(1.0, float(want_in[l].grad.abs().max()))
        np.testing.assert_allclose(a[l].cpu().numpy(), c[l].cpu().numpy(), rtol=0, atol=2e-5 * scale)   # atomics: order-dependent
        np.testing.assert_allclose(a[l].cpu().numpy(), want_in[l].grad.float().numpy(), rtol=0, atol=3e-6 * scale)
        empty = [n for n, cnt in enumerate(counts) if cnt == 0 and not ((rois[:, 0] == n).any())]
        for n in empty:
            assert float(a[l][n].abs().sum()) == 0.0


def _tc_case(gen, N, H, W, counts, C=128):
    """feature levels + a ROI table with `counts[n]` boxes for frame n, rows shuffled, plus edge-case boxes"""
    feats = _feats(gen, N, C, H, W)
    rows = [torch.cat([torch.full((c, 1), float(n)), _boxes(gen, c, H, W)], 1) for n, c in enumerate(counts) if c > 0]
    rois = torch.cat(rows, 0)
    edge = torch.tensor([[0, -30.0, -10.0, 20.0, 40.0], [N - 1, W - 8.0, H - 6.0, W + 22.0, H + 24.0],   # sticking out
                         [0, 100.0, 100.0, 100.0, 100.0], [0, 0.0, 0.0, W - 1.0, H - 1.0],              # point, whole image
                         [N - 1, -500.0, -500.0, -400.0, -400.0], [0, 3.3, 4.4, 9.9, 8.8],              # fully outside, tiny
                         [-1, 10.0, 10.0, 50.0, 50.0], [N, 10.0, 10.0, 50.0, 50.0]])                   # frames that do not exist
    rois = torch.cat([rois, edge], 0)
    return feats, rois[torch.randperm(rois.shape[0], generator=gen)]


@pytest.mark.parametrize("N,H,W,counts", [
    (3, 256, 448, [50, 0, 9]),             # 64x112 / 32x56 / 16x28 rows on the tensor cores, the 8x14 level in linear mode
    (2, 255, 448, [130, 70]),              # the scripts' 255x448; 3 + 2 groups of <= 64 ROIs per frame
    (2, 128, 192, [20, 5]),                # level 3 is 4x6: linear mode
    (1, 160, 224, [33]),                   # 10x14 and 5x7 cannot go through TMA: those two levels use the gather kernel
    (1, 64, 1024, [12]),                   # 256-wide level 0: two 128-element x segments
    (9, 64, 96, [7] * 9),
])
def test_roi_mean_pool_tensor_core_path(N, H, W, counts):
    """K5-TC (csrc/roi_pool_tc.cu) against the ROIAlign stand-in oracle and against the gather kernel; frames with
    several ROI groups, empty frames, shuffled rows, boxes sticking out of / outside the image, invalid frame ids."""
    gen = torch.Generator().manual_seed(100 + N + H)
    feats, rois = _tc_case(gen, N, H, W, counts)
    dfeats, drois = [f.to(DEV) for f in feats], rois.to(DEV)
    got = ops.roi_mean_pool(dfeats, drois, impl="tc")
    simt = ops.roi_mean_pool(dfeats, drois, impl="simt")
    ok = (rois[:, 0] >= 0) & (rois[:, 0] < N)
    want = torch.zeros(rois.shape[0], 512)
    want[ok] = orc.roi_mean_pool(feats, rois[ok])
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(got.cpu().numpy(), simt.cpu().numpy(), rtol=0, atol=5e-6)
    assert float(got[~ok.to(DEV)].abs().sum()) == 0.0
    assert torch.equal(got, ops.roi_mean_pool(dfeats, drois, impl="auto"))            # auto == tc inside the envelope
    # the same box pools to the same bits whatever table it arrives in (row order, neighbours, table size)
    sub = torch.cat([drois[5:12], drois[:3]], 0)
    assert torch.equal(ops.roi_mean_pool(dfeats, sub, impl="tc"), torch.cat([got[5:12], got[:3]], 0))
    assert torch.equal(got, ops.roi_mean_pool(dfeats, drois, impl="tc"))              # run to run


def test_roi_mean_pool_tensor_core_envelope():
    gen = torch.Generator().manual_seed(5)
    feats = [f.to(DEV) for f in _feats(gen, 1, 128, 168, 216)]                          # 42x54, 21x27, 10x13, 5x6: no level fits
    rois = torch.tensor([[0, 10.0, 10.0, 90.0, 120.0]], device=DEV)
    with pytest.raises(RuntimeError):
        ops.roi_mean_pool(feats, rois, impl="tc")
    a = ops.roi_mean_pool(feats, rois, impl="auto")
    assert torch.equal(a, ops.roi_mean_pool(feats, rois, impl="simt"))
    feats24 = [f.to(DEV) for f in _feats(gen, 1, 24, 256, 448)]                         # C != 128 -> gather kernel
    with pytest.raises(RuntimeError):
        ops.roi_mean_pool(feats24, rois, impl="tc")


def test_feature_extractor_module_layout():
    gen = torch.Generator().manual_seed(8)
    feats = [f.to(DEV) for f in _feats(gen, 2, 128, 128, 192)]
    props = [BoxList(_boxes(gen, 5, 128, 192)).to(DEV), BoxList(_boxes(gen, 3, 128, 192)).to(DEV)]
    fe = make_roi_mask_feature_extractor()
    out = fe(tuple(feats), props)
    assert out.shape == (8, 512)
    rois = fe.convert_to_roi_format(props)
    assert rois.shape == (8, 5) and rois[:5, 0].eq(0).all() and rois[5:, 0].eq(1).all()
    want = orc.roi_mean_pool([f.cpu() for f in feats], rois.cpu())
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=0, atol=1e-5)


@pytest.mark.parametrize("is_test", [1, 0])
def test_batched_container_equals_per_video_loop(is_test):
    B, P, Fm, H, W, C = 4, 12, 5, 64, 96, 16
    gen = torch.Generator().manual_seed(21)
    pr = make_problems(B, P, Fm, H, W, 4 * C, seed=31, with_targets=True)
    feats = _feats(gen, B, C, H, W)
    n_prop = [12, 7, 12, 3]
    valid = torch.tensor([[1, 1, 1, 0, 0], [1, 0, 0, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]], dtype=torch.float32)
    boxes = [_boxes(gen, n, H, W) for n in n_prop]
    tboxes = [_boxes(gen, Fm, H, W) for _ in range(B)]
    cfg = default_cfg(10, 5)

    def boxlists(dev):
        out = []
        for b in range(B):
            bl = BoxList(boxes[b])
            bl.add_field('mask', pr.prop_mask[b, :n_prop[b]].unsqueeze(1))
            bl.add_field('scores', pr.prop_score[b, :n_prop[b]])
            out.append(bl.to(dev))
        return out

    # ---- oracle: pooled features via torchvision ROIAlign, then the per-video loop ------------------------
    rois = torch.cat([torch.cat([torch.full((n_prop[b], 1), float(b)), boxes[b]], 1) for b in range(B)], 0)
    pooled = orc.roi_mean_pool(feats, rois).split(n_prop, 0)
    trois = torch.cat([torch.cat([torch.full((Fm, 1), float(b)), tboxes[b]], 1) for b in range(B)], 0)
    tfeat = orc.roi_mean_pool(feats, trois).split([Fm] * B, 0)
    want_out, want_loss, want_last = orc.dmm_container_forward(
        cfg, is_test, list(pooled), [pr.prop_mask[b, :n_prop[b]] for b in range(B)],
        [pr.prop_score[b, :n_prop[b]] for b in range(B)], list(tfeat), pr.tmpl_mask, valid,
        None if is_test else pr.targets)

    # ---- product: batched container ----------------------------------------------------------------------
    model = DMM_Model(cfg, is_test=is_test).to(DEV)
    dfeats = tuple(f.to(DEV) for f in feats)
    tplt = model.fill_template_dict(None, [BoxList(tboxes[b]).to(DEV) for b in range(B)],
                                    {'backbone_feature': dfeats, 'refine_input_feat': dfeats}, None, valid)
    if is_test:
        with torch.no_grad():
            out, _, loss, last = model.inference({'args': None, 'shape': (H, W), 'extra_frame': [0] * B, 'valid': valid.to(DEV)},
                                                 boxlists(DEV), dfeats, pr.tmpl_mask.to(DEV), tplt)
        assert loss == []
    else:
        out, _, loss, last = model(None, boxlists(DEV), dfeats, pr.tmpl_mask.to(DEV), tplt, valid.to(DEV), pr.targets.to(DEV))
        assert len(loss) == B
        k = 0
        for b in range(B):
            if valid[b].sum() == 0:
                assert float(loss[b]) == 0.0
            np.testing.assert_allclose(float(loss[b]), float(want_loss[b]), rtol=0, atol=1e-5)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want_out.numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(last.detach().cpu().numpy(), want_last.numpy(), rtol=0, atol=1e-4)


def test_ragged_pointer_table_equals_stacked_batch():
    """Per-video proposal tensors through the device pointer table == the padded [B,P,H,W] batch (K1 and K4)."""
    B, P, O, H, W, D = 4, 14, 3, 40, 52, 32
    pr = make_problems(B, P, O, H, W, D, seed=77).to(DEV)
    counts = [14, 9, 1, 6]
    plist = [pr.prop_mask[b, :n].clone() for b, n in enumerate(counts)]          # separate allocations
    n_prop = torch.tensor(counts, device=DEV, dtype=torch.int32)
    a = ops.mask_iou_pairwise(pr.prop_mask, pr.tmpl_mask, n_prop=n_prop)["iou"]
    b_ = ops.mask_iou_pairwise(plist, pr.tmpl_mask)["iou"]
    assert torch.equal(a, b_)
    kw = dict(max_iter=20, proj_iter=5, lr=0.1, score_weight=0.3, is_test=True)
    with torch.no_grad():
        ref = ops.match_batch(pr.prop_feat, pr.prop_mask, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score, n_prop=n_prop, **kw)
        got = ops.match_batch(pr.prop_feat, plist, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score, **kw)
    for k in ("full_outmask", "match_score", "det_score", "R"):
        assert torch.equal(ref[k], got[k]), k
    # training direction: gradient w.r.t. the assignment through the pointer-table backward
    Bm = ref["Bmat"].clone().requires_grad_(True)
    Bm2 = ref["Bmat"].clone().requires_grad_(True)
    w = torch.rand(B, O, H * W, device=DEV)
    (ops.assign_apply(Bm, pr.prop_mask, ref["logic"], n_prop=n_prop) * w).sum().backward()
    (ops.assign_apply(Bm2, plist, ref["logic"]) * w).sum().backward()
    assert torch.allclose(Bm.grad, Bm2.grad, rtol=1e-5, atol=1e-5)


def test_synthetic_clip_loop_runs_sequential_frames():
    """examples/synthetic_clip_eval.py: the eval-shaped loop (templates of frame t+1 = matched masks of frame t)."""
    import argparse
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "synthetic_clip_eval.py")
    spec = importlib.util.spec_from_file_location("synthetic_clip_eval", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rate = mod.run(argparse.Namespace(clips=3, frames=4, proposals=9, objects=3, size=[64, 96], lazy=False))
    assert rate > 0
    assert mod.run(argparse.Namespace(clips=3, frames=4, proposals=9, objects=3, size=[64, 96], lazy=True)) > 0


def test_synthetic_train_step_runs_and_produces_gradients():
    """examples/synthetic_train_step.py (BASELINE configs[4] shape, shrunk): encoder -> K8 paste -> K5 -> training-mode
    container (K2/K1/K3/K4 with backward) -> K6 pyramid (backward) -> decoder -> loss -> backward -> Adam."""
    import argparse
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "synthetic_train_step.py")
    spec = importlib.util.spec_from_file_location("synthetic_train_step", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hist = mod.run(argparse.Namespace(steps=3, clips=2, frames=3, objects=2, proposals=7, arch="resnet18", size=[64, 96]))
    assert len(hist) == 3 and all(h[0] == h[0] for h in hist)


def test_lazy_pipeline_is_bit_identical_to_paste_nms_match():
    """DMM_Model.inference_lazy (bits-only paste, device-side keep table, packed K1, fused paste+apply) ==
    Masker -> filter_results -> DMM_Model.inference, frame after frame (the outputs feed the next frame's templates)."""
    from dmm_net_b200.utils.boxlist_ops import filter_results
    from dmm_net_b200.utils.masker import Masker
    gen = torch.Generator(device=DEV).manual_seed(21)
    B, F, H, W, C, n_det, P = 3, 4, 64, 96, 16, [23, 17, 30], 12
    model = DMM_Model(default_cfg(20, 5), is_test=1).to(DEV)
    valid = torch.tensor([[1, 1, 1, 0], [1, 0, 0, 0], [1, 1, 1, 1]], device=DEV).float()
    feats = lambda: tuple(torch.randn(B, C, H // s, W // s, generator=gen, device=DEV) for s in (4, 8, 16, 32))

    def boxes(n):
        xy = torch.rand(n, 2, generator=gen, device=DEV) * torch.tensor([W * 0.6, H * 0.6], device=DEV)
        wh = torch.rand(n, 2, generator=gen, device=DEV) * torch.tensor([W * 0.4, H * 0.4], device=DEV) + 4
        return torch.cat([xy, (xy + wh).clamp(max=min(H, W) - 1)], 1)

    f0 = feats()
    tb = [boxes(F) for _ in range(B)]
    tplt = model.fill_template_dict(None, [BoxList(b) for b in tb], {"backbone_feature": f0, "refine_input_feat": f0}, None, valid)
    masker = Masker(0.5, 1)
    m0, _ = masker([torch.ones(F, 1, 28, 28, device=DEV)] * B, [BoxList(b, (W, H)) for b in tb])
    last_a = last_b = torch.stack([m.squeeze(1) for m in m0], 0) * valid[:, :, None, None]
    infos = {"args": None, "shape": (H, W), "extra_frame": [0, 0, 0], "valid": valid}
    for t in range(3):
        fb = feats()
        dets = []
        for b in range(B):
            d = BoxList(boxes(n_det[b]), (W, H))
            d.add_field("mask", torch.sigmoid(3 * torch.randn(n_det[b], 1, 28, 28, generator=gen, device=DEV) + 1))
            d.add_field("scores", (torch.rand(n_det[b], generator=gen, device=DEV) * 10).round() / 10)   # tied scores
            dets.append(d)
        # reference-shaped path: paste everything, NMS on the tight boxes, match
        pasted, tight = masker([d.get_field("mask") for d in dets], dets)
        props = []
        for b, d in enumerate(dets):
            bl = BoxList(tight[b].float(), (W, H))
            bl.add_field("mask", pasted[b])
            bl.add_field("scores", d.get_field("scores"))
            props.append(bl)
        props = filter_results(props, nms_thresh=0.8, max_proposals=P)
        with torch.no_grad():
            out_a, _, _, last_a = model.inference(infos, props, fb, last_a, tplt)
            out_b, _, _, last_b, (src_index, n_prop) = model.inference_lazy(infos, dets, fb, last_b, tplt, 0.8, P)
        assert n_prop.tolist() == [len(p) for p in props]
        assert torch.equal(out_a, out_b), t
        assert torch.equal(last_a, last_b), t
        assert float(out_a.abs().sum()) > 0


def test_lazy_training_equals_paste_all_training():
    """DMM_Model.forward_lazy (bits-only paste, packed K1 with the targets as second set, K10 with its backward) ==
    Masker -> filter_results -> DMM_Model.forward in training mode: same outputs and match loss, same gradients into the
    backbone feature maps (K10's backward against K4's), without materialising the pasted proposal masks."""
    from dmm_net_b200.utils.boxlist_ops import filter_results
    from dmm_net_b200.utils.masker import Masker
    gen = torch.Generator(device=DEV).manual_seed(33)
    B, F, H, W, C, n_det, P = 3, 4, 64, 96, 16, [23, 17, 30], 12
    model = DMM_Model(default_cfg(10, 5), is_test=0).to(DEV)
    valid = torch.tensor([[1, 1, 1, 0], [1, 0, 0, 0], [1, 1, 1, 1]], device=DEV).float()

    def boxes(n):
        xy = torch.rand(n, 2, generator=gen, device=DEV) * torch.tensor([W * 0.6, H * 0.6], device=DEV)
        wh = torch.rand(n, 2, generator=gen, device=DEV) * torch.tensor([W * 0.4, H * 0.4], device=DEV) + 4
        return torch.cat([xy, (xy + wh).clamp(max=min(H, W) - 1)], 1)

    base = [torch.randn(B, C, H // s, W // s, generator=gen, device=DEV) for s in (4, 8, 16, 32)]
    tb = [boxes(F) for _ in range(B)]
    masker = Masker(0.5, 1)
    m0, _ = masker([torch.ones(F, 1, 28, 28, device=DEV)] * B, [BoxList(b, (W, H)) for b in tb])
    last = torch.stack([m.squeeze(1) for m in m0], 0) * valid[:, :, None, None]
    targets = (torch.roll(last, 3, dims=3) > 0.5).float()
    dets = []
    for b in range(B):
        d = BoxList(boxes(n_det[b]), (W, H))
        d.add_field("mask", torch.sigmoid(3 * torch.randn(n_det[b], 1, 28, 28, generator=gen, device=DEV) + 1))
        d.add_field("scores", torch.rand(n_det[b], generator=gen, device=DEV))
        dets.append(d)
    w_out = torch.rand(B, F, H, W, generator=gen, device=DEV)

    def run(lazy):
        feats = tuple(f.clone().requires_grad_(True) for f in base)
        tplt = model.fill_template_dict(None, [BoxList(b, (W, H)) for b in tb], {"backbone_feature": feats, "refine_input_feat": feats},
                                        None, valid)
        if lazy:
            out, _, loss, lastn, _ = model.forward_lazy(None, dets, feats, last, tplt, valid, targets, 0.8, P)
        else:
            pasted, tight = masker([d.get_field("mask") for d in dets], dets)
            props = []
            for b, d in enumerate(dets):
                bl = BoxList(tight[b].float(), (W, H))
                bl.add_field("mask", pasted[b])
                bl.add_field("scores", d.get_field("scores"))
                props.append(bl)
            props = filter_results(props, nms_thresh=0.8, max_proposals=P)
            out, _, loss, lastn = model(None, props, feats, last, tplt, valid, targets)
        ((out * w_out).sum() + 3.0 * sum(loss)).backward()
        return out.detach(), torch.stack([x.detach() for x in loss]), lastn.detach(), [f.grad for f in feats]

    out_a, loss_a, last_a, g_a = run(False)
    out_b, loss_b, last_b, g_b = run(True)
    np.testing.assert_allclose(out_b.cpu().numpy(), out_a.cpu().numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(last_b.cpu().numpy(), last_a.cpu().numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(loss_b.cpu().numpy(), loss_a.cpu().numpy(), rtol=0, atol=1e-6)
    assert float(out_a.abs().sum()) > 0 and float(loss_a.abs().sum()) > 0
    for ga, gb in zip(g_a, g_b):
        scale = max(1.0, float(ga.abs().max()))
        np.testing.assert_allclose(gb.cpu().numpy(), ga.cpu().numpy(), rtol=0, atol=2e-5 * scale)
    assert any(float(g.abs().sum()) > 0 for g in g_a)
