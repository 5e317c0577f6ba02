"""K2 on the tensor cores (tcgen05 3xTF32, cosine_tc.cu) against the fp32 FFMA kernel, the CPU oracle and an fp64
reference.  Bar: 1e-4 (BASELINE.json north_star); asserted here: 5e-6 against fp64."""
import pytest
import torch

from dmm_net_b200 import ops
from oracle import match_oracle as orc

pytestmark = pytest.mark.gpu


def ref64(q, k, eps=1e-8):
    q, k = q.double(), k.double()
    qn = q.norm(dim=-1).clamp_min(eps)[..., :, None]
    kn = k.norm(dim=-1).clamp_min(eps)[..., None, :]
    return (q @ k.transpose(-1, -2)) / (qn * kn)


def feats(B, P, O, D, seed):
    g = torch.Generator().manual_seed(seed)
    k = torch.randn(B, P, D, generator=g)
    q = k[:, torch.arange(O) % P] + 0.3 * torch.randn(B, O, D, generator=g)      # planted matches: cos near 1 and near 0
    return q.cuda(), k.cuda()


@pytest.mark.parametrize("B,P,O,D", [(1, 50, 10, 512), (2, 50, 10, 512), (5, 64, 16, 512), (9, 13, 5, 64), (3, 1, 1, 32),
                                      (130, 50, 10, 128), (7, 33, 7, 260), (4, 50, 10, 36), (301, 50, 10, 512)])
def test_tc_matches_fp64_and_simt(B, P, O, D):
    q, k = feats(B, P, O, D, B * 1000 + D)
    if B > 2:
        k[1, 0].zero_()                                    # zero vector: norm clamped at eps, cos = 0
        k[2] *= 1e3                                        # scale invariance
        q[0] *= 1e-3
    tc = ops.cosine_pairwise(q[:, None], k, impl="tc")
    simt = ops.cosine_pairwise(q[:, None], k, impl="simt")
    auto = ops.cosine_pairwise(q[:, None], k)
    want = ref64(q, k)
    assert (tc.double() - want).abs().max().item() <= 5e-6
    assert (simt.double() - want).abs().max().item() <= 2e-6
    assert (tc - simt).abs().max().item() <= 5e-6
    assert torch.equal(auto, tc)                            # the default inside the envelope is the tensor-core kernel
    if B > 2:
        assert (tc[1, :, 0] == 0).all()
    oracle = torch.stack([orc.cosine_scores(q[b].cpu(), k[b].cpu()) for b in range(min(B, 4))])
    assert (tc[:4].cpu() - oracle).abs().max().item() <= 5e-6


def test_tc_ragged_counts_define_padding_as_zero():
    B, P, O, D = 6, 50, 10, 256
    q, k = feats(B, P, O, D, 3)
    n_prop = torch.tensor([50, 1, 0, 37, 50, 12], dtype=torch.int32).cuda()
    n_tmpl = torch.tensor([10, 3, 5, 0, 1, 10], dtype=torch.int32).cuda()
    tc = ops.cosine_pairwise(q[:, None], k, n_prop, n_tmpl, impl="tc")
    simt = ops.cosine_pairwise(q[:, None], k, n_prop, n_tmpl, impl="simt")
    assert (tc - simt).abs().max().item() <= 5e-6
    for b in range(B):
        assert (tc[b, int(n_tmpl[b]):] == 0).all() and (tc[b, :, int(n_prop[b]):] == 0).all()


def test_tc_envelope_and_fallback():
    q, k = feats(3, 70, 10, 64, 1)                         # P > 64: outside the envelope
    with pytest.raises(RuntimeError):
        ops.cosine_pairwise(q[:, None], k, impl="tc")
    auto = ops.cosine_pairwise(q[:, None], k)               # auto falls back to the FFMA kernel
    assert (auto.double() - ref64(q, k)).abs().max().item() <= 2e-6
    q, k = feats(2, 20, 4, 66, 2)                          # D % 4 != 0
    assert (ops.cosine_pairwise(q[:, None], k).double() - ref64(q, k)).abs().max().item() <= 2e-6
    q, k = feats(2, 20, 4, 64, 3)                          # two template-feature sets: mean of the two cosines, FFMA kernel
    q2 = torch.stack([q, q.flip(1)], 1)
    got = ops.cosine_pairwise(q2, k)
    want = 0.5 * (ref64(q, k) + ref64(q.flip(1), k))
    assert (got.double() - want).abs().max().item() <= 2e-6
    qm = torch.randn(2, 1, 4, 65, device="cuda")[..., 1:]   # misaligned rows -> contiguous copy keeps the envelope
    assert (ops.cosine_pairwise(qm, k).double() - ref64(qm[:, 0], k)).abs().max().item() <= 5e-6


def test_tc_forward_feeds_the_simt_backward():
    """training: forward on the tensor cores, backward (cosine.cu) reuses the forward's cosines"""
    q, k = feats(3, 20, 5, 128, 7)
    q1, k1 = q[:, None].clone().requires_grad_(True), k.clone().requires_grad_(True)
    w = torch.rand(3, 5, 20, device="cuda")
    (ops.cosine_pairwise(q1, k1, impl="tc") * w).sum().backward()
    q2, k2 = q.double().requires_grad_(True), k.double().requires_grad_(True)
    (ref64(q2, k2) * w.double()).sum().backward()
    assert (q1.grad[:, 0].double() - q2.grad).abs().max().item() <= 1e-5
    assert (k1.grad.double() - k2.grad).abs().max().item() <= 1e-5
