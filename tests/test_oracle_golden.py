"""The CPU oracle against the reference's own outputs (tests/golden, made by oracle/make_golden.py)
and against the one known answer the reference carries (relax_match.py:108-119)."""
import numpy as np
import pytest
import torch

from conftest import T, golden_names, load_golden
from oracle import match_oracle as orc
from dmm_net_b200.synth import default_cfg

torch.set_num_threads(1)


@pytest.mark.parametrize("name", golden_names("solver_"))
def test_solver_matches_reference(name):
    g = load_golden(name)
    C = T(g["C"])
    mi, pi = [int(v) for v in g["params"]]
    X, cost, X_list, _ = orc.relax_solve(C, mi, pi, float(g["lr"]))
    assert len(X_list) == int(g["n_list"])                       # both early exits fire at the same step
    assert torch.equal(X_list[0], T(g["X0"]))                   # greedy start bit-exact
    R = sum(X_list) / len(X_list)
    np.testing.assert_array_equal(R.numpy(), g["R"])            # same ops, same order -> bit-exact
    np.testing.assert_array_equal(X.numpy(), g["X"])
    np.testing.assert_array_equal(np.array(cost, np.float64), g["cost"])


def test_solver_known_answer_equals_hungarian():
    g = load_golden("solver_known_3x3")
    C = T(g["C"])
    X, _, X_list, _ = orc.relax_solve(C, 100, 100, 0.1)
    assert len(X_list) == 58                                    # outer exit after 57 gradient steps
    want = torch.tensor([[0., 1, 0], [1, 0, 0], [0, 0, 1]])
    assert torch.equal(orc.hungarian_onehot(C), want)
    assert (X - want).abs().max() < 1e-3
    assert torch.equal(X.argmax(1), want.argmax(1))


def test_solver_invariants():
    gen = torch.Generator().manual_seed(3)
    C = -torch.rand(6, 17, generator=gen)
    X, _, X_list, _ = orc.relax_solve(C, 30, 50, 0.1)
    assert torch.equal(X_list[0].sum(1), torch.ones(6)) and set(X_list[0].unique().tolist()) <= {0.0, 1.0}
    np.testing.assert_array_equal(X_list[1].numpy(), (X_list[0] - 0.1 * C).numpy())   # pre-projection iterate
    # Dykstra is only run for a bounded number of sweeps: feasibility holds to its tolerance
    assert (X.sum(1) - 1).abs().max() < 1e-4 and X.min() > -1e-2 and X.sum(0).max() < 1 + 1e-2


def test_rowwise_iou_bit_exact():
    g = load_golden("iou_rows")
    got = orc.rowwise_binary_iou(T(g["a"]), T(g["b"]))
    np.testing.assert_array_equal(got.numpy(), g["iou"])
    assert got[3] == 0                                           # empty vs empty -> 0 / 1e-6


def test_cosine_bit_exact():
    g = load_golden("cosine")
    np.testing.assert_array_equal(orc.cosine_scores(T(g["q"]), T(g["k"])).numpy(), g["cos"])


@pytest.mark.parametrize("name", golden_names("layer_"))
def test_layer_matches_reference(name):
    g = load_golden(name)
    P, O, H, W, D, mi, pi, is_test = [int(v) for v in g["meta"]]
    cfg = default_cfg(mi, pi, float(g["lr"]), float(g["score_weight"]))
    pf = T(g["prop_feat"]).requires_grad_(not is_test)
    tf = T(g["tmpl_feat"]).requires_grad_(not is_test)
    sc = T(g["prop_score"]).requires_grad_(not is_test)
    tg = T(g["targets"]) if "targets" in g else None
    pm, tm = T(g["prop_mask"]), T(g["tmpl_mask"])
    for expand in (True, False):
        np.testing.assert_array_equal(orc.pairwise_binary_iou(pm.view(P, -1), tm.view(O, -1), expand).numpy(), g["iou"])
    sim, _ = orc.cost_matrix(pf, pm, [tf], tm, cfg["score_weight"], tg)
    np.testing.assert_array_equal(sim.detach().numpy(), g["sim"])
    full, ms, ds, full2, loss = orc.match_layer_forward(cfg, is_test, pf, pm, [tf], tm, sc, tg)
    assert full is full2
    np.testing.assert_array_equal(full.detach().numpy(), g["full_outmask"])
    np.testing.assert_array_equal(ms.detach().numpy(), g["match_score"])
    np.testing.assert_array_equal(ds.detach().numpy(), g["det_score"])
    if "cost_loss" in g:
        np.testing.assert_array_equal(loss["cost_loss"].detach().numpy(), g["cost_loss"])
    if not is_test:
        total = (full * T(g["w_mask"])).sum() + (ms * T(g["w_ms"])).sum() + (ds * T(g["w_ds"])).sum()
        if "cost_loss" in loss:
            total = total + 3.0 * loss["cost_loss"]
        total.backward()
        np.testing.assert_allclose(pf.grad.numpy(), g["g_prop_feat"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(tf.grad.numpy(), g["g_tmpl_feat"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(sc.grad.numpy(), g["g_prop_score"], rtol=1e-5, atol=1e-6)


def test_roi_pool_separable_equals_roialign():
    gen = torch.Generator().manual_seed(11)
    feats = [torch.randn(2, 6, 64 // s, 96 // s, generator=gen) for s in (1, 2, 4, 8)]   # image 256 x 384
    rois = torch.tensor([[0, 10.3, 20.7, 120.2, 200.9], [1, 0, 0, 383, 255], [0, 300.5, 100.25, 310.0, 104.0],
                         [1, 50, 60, 50, 60], [0, -20.0, -8.0, 40.0, 30.0], [1, 350.0, 200.0, 420.0, 300.0]])
    a = orc.roi_mean_pool(feats, rois)
    b = orc.roi_mean_pool_separable(feats, rois)
    assert a.shape == (6, 24)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=0, atol=2e-6)
