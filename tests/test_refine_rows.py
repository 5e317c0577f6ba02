"""Rows after the layer: decoder mask-input pyramid (K6), merged label map (K7), hard-IoU mean.

CPU part: the oracle (oracle/refine_oracle.py) against the golden vectors that oracle/make_golden_refine.py produced by
executing the reference's own source lines.  GPU part: the CUDA kernels through the C ABI against those vectors and
against the oracle on fresh inputs -- max / argmax are exact, so everything here is compared with array equality."""
import numpy as np
import pytest
import torch

from conftest import T, golden_names, load_golden
from oracle import refine_oracle as rorc

PYR = golden_names("refine_pyramid_")
LAB = golden_names("refine_labels_")


def test_golden_present():
    assert len(PYR) >= 3 and len(LAB) >= 2


# ---------------------------------------------------------------------------------------------------------
# oracle vs the reference's lines (CPU)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", PYR)
def test_oracle_pyramid_matches_reference_lines(name):
    g = load_golden(name)
    B, O, H, W, L = (int(v) for v in g["meta"])
    prev, ref, init = (T(g[k]).requires_grad_(True) for k in ("prev", "ref", "init"))
    levels = rorc.mask_pyramid(prev.view(B, O, H, W), ref.view(B, O, H, W), init, L)
    for k in range(L):
        assert np.array_equal(levels[k].detach().numpy(), g[f"level{k}"]), k
    sum((lv * T(g[f"w{k}"])).sum() for k, lv in enumerate(levels)).backward()
    for n, t in (("g_prev", prev), ("g_ref", ref), ("g_init", init)):
        assert np.array_equal(t.grad.numpy().reshape(g[n].shape), g[n]), n


@pytest.mark.parametrize("name", LAB)
def test_oracle_labels_match_reference_lines(name):
    g = load_golden(name)
    lab = rorc.merged_labels(T(g["outs"]), T(g["n_valid"]))
    assert np.array_equal(lab.numpy(), g["label"])


def test_chained_ceil_pools_equal_clipped_windows():
    """the identity K6 is built on: 1+k chained ceil-mode 2x2 pools == max over the clipped (4<<k)-wide window"""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 1, 45, 83, generator=g)
    lv = rorc.mask_pyramid(x, x, x, 4)
    for k in range(4):
        win = 4 << k
        want = torch.nn.functional.max_pool2d(x, win, win, ceil_mode=True)
        assert torch.equal(lv[k][0, :, 0], want[:, 0]), k


# ---------------------------------------------------------------------------------------------------------
# CUDA kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", PYR)
def test_k6_pyramid_golden_fwd_bwd(name):
    from dmm_net_b200 import ops
    g = load_golden(name)
    B, O, H, W, L = (int(v) for v in g["meta"])
    prev, ref, init = (T(g[k], "cuda").requires_grad_(True) for k in ("prev", "ref", "init"))
    levels = ops.mask_pyramid(prev, ref, init, L)
    assert len(levels) == L
    for k in range(L):
        assert np.array_equal(levels[k].detach().cpu().numpy(), g[f"level{k}"]), (name, k)
    sum((lv * T(g[f"w{k}"], "cuda")).sum() for k, lv in enumerate(levels)).backward()
    for n, t in (("g_prev", prev), ("g_ref", ref), ("g_init", init)):
        assert np.array_equal(t.grad.cpu().numpy().reshape(g[n].shape), g[n]), (name, n)


@pytest.mark.gpu
@pytest.mark.parametrize("B,O,H,W,L", [(2, 3, 256, 448, 4), (1, 2, 255, 447, 4), (3, 1, 64, 64, 5), (1, 1, 5, 3, 3),
                                         (2, 2, 130, 200, 1)])
def test_k6_pyramid_vs_oracle(B, O, H, W, L):
    from dmm_net_b200 import ops
    gen = torch.Generator().manual_seed(H * 7 + W)
    mk = lambda: ((torch.rand(B, O, H, W, generator=gen) * 16).round() / 16)      # quantised: plenty of tied maxima
    prev, ref, init = mk(), mk(), mk()
    init[:, :, : H // 2] = 0                                                       # exact-zero regions (ties everywhere)
    with_nan = H == 255
    if with_nan:
        prev[0, 0, 3, 1] = float("nan")                                            # NaN propagates through every level
    cpu = [t.clone().requires_grad_(True) for t in (prev, ref, init)]
    want = rorc.mask_pyramid(*cpu, L)
    dev = [t.cuda().requires_grad_(True) for t in (prev, ref, init)]
    got = ops.mask_pyramid(dev[0].view(B, O, H * W), dev[1], dev[2], L)           # prev as [B,O,HW], like the trainer holds it
    ws = [torch.rand(w.shape, generator=gen) for w in want]
    for k in range(L):
        assert got[k].shape == want[k].shape
        assert np.array_equal(got[k].detach().cpu().numpy(), want[k].detach().numpy(), equal_nan=True), k
    if with_nan:
        return                                                                     # NaN case: forward only
    sum((a * w).sum() for a, w in zip(want, ws)).backward()
    sum((a * w.cuda()).sum() for a, w in zip(got, ws)).backward()
    for a, b in zip(dev, cpu):
        assert np.array_equal(a.grad.cpu().numpy().reshape(b.grad.shape), b.grad.numpy())


@pytest.mark.gpu
def test_k6_pyramid_batch_strided_views_and_partial_grads():
    """prev_mask / ref_mask arrive as views into larger tensors (targets[0][:,:,:-1]) and often need no gradient"""
    from dmm_net_b200 import ops
    gen = torch.Generator().manual_seed(5)
    B, O, H, W = 2, 3, 72, 100
    big = torch.rand(B, O + 2, H, W, generator=gen).cuda()
    ref = big[:, :O]                                                               # batch stride (O+2)*H*W
    prev = torch.rand(B, O, H * W, generator=gen).cuda()
    init = torch.rand(B, O, H, W, generator=gen).cuda().requires_grad_(True)
    got = ops.mask_pyramid(prev, ref, init, 4)
    want = rorc.mask_pyramid(prev.cpu().view(B, O, H, W), ref.cpu(), init.detach().cpu(), 4)
    for a, b in zip(got, want):
        assert torch.equal(a.detach().cpu(), b)
    got[2].sum().backward()
    ci = init.detach().cpu().requires_grad_(True)
    rorc.mask_pyramid(prev.cpu().view(B, O, H, W), ref.cpu(), ci, 4)[2].sum().backward()
    assert torch.equal(init.grad.cpu(), ci.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("name", LAB)
def test_k7_labels_golden(name):
    from dmm_net_b200 import ops
    g = load_golden(name)
    lab = ops.merge_labels(T(g["outs"], "cuda"), T(g["n_valid"], "cuda"))
    assert np.array_equal(lab.cpu().numpy(), g["label"])


@pytest.mark.gpu
@pytest.mark.parametrize("B,O,HW", [(4, 5, 256 * 448), (2, 3, 255 * 447), (3, 1, 17), (1, 7, 4096)])
def test_k7_labels_vs_oracle(B, O, HW):
    from dmm_net_b200 import ops
    gen = torch.Generator().manual_seed(HW)
    outs = (torch.rand(B, O, HW, generator=gen) * 8).round() / 8                   # ties between objects and with 1-max
    n_valid = torch.randint(0, O + 1, (B,), generator=gen)
    lab = ops.merge_labels(outs.cuda(), n_valid.cuda())
    assert torch.equal(lab.cpu(), rorc.merged_labels(outs, n_valid))
    full = ops.merge_labels(outs.cuda().view(B, O, 1, HW))                         # no counts: all O objects valid
    assert torch.equal(full.cpu(), rorc.merged_labels(outs, torch.full((B,), O)))


@pytest.mark.gpu
def test_hard_iou_mean_vs_oracle():
    from dmm_net_b200 import ops
    gen = torch.Generator().manual_seed(9)
    B, O, HW = 3, 4, 64 * 112
    y = (torch.rand(B, O, HW, generator=gen) > 0.6).float()
    pred = torch.rand(B, O, HW, generator=gen)
    valid = torch.tensor([[1, 1, 0, 0], [1, 0, 0, 0], [1, 1, 1, 1]])
    got = ops.hard_iou_mean(y.cuda(), pred.cuda(), valid.cuda())
    want = rorc.hard_iou_mean(y, pred, valid)
    assert abs(got.item() - want.item()) <= 1e-6
    none = ops.hard_iou_mean(y.cuda(), pred.cuda(), torch.zeros_like(valid).cuda())
    assert none.item() == 0.0
