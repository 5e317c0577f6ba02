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


# ---------------------------------------------------------------------------------------------------------
# goldens of oracle/make_golden_container.py: the reference's DMM_Model container, algo='hun', the headline size
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("container_"))
def test_container_matches_reference(name):
    """oracle.dmm_container_forward == the unmodified reference DMM_Model.forward / .inference (dmm_model.py:48-158),
    given the same pooled features (the reference's own FeatureExtractor over the stand-in ROIAlign)."""
    g = load_golden(name)
    B, P, Fm, H, W, C, mi, pi, is_test = [int(v) for v in g["meta"]]
    n_prop = [int(v) for v in g["n_prop"]]
    cfg = default_cfg(mi, pi)
    feats = [T(g[f"feat{l}"]) for l in range(4)]
    rois = torch.cat([torch.cat([torch.full((n_prop[b], 1), float(b)), T(g[f"boxes{b}"])], 1) for b in range(B)], 0)
    pooled = orc.roi_mean_pool(feats, rois)
    np.testing.assert_array_equal(pooled.numpy(), g["prop_pooled"])          # same stand-in ROIAlign arithmetic
    trois = torch.cat([torch.cat([torch.full((Fm, 1), float(b)), T(g[f"tboxes{b}"])], 1) for b in range(B)], 0)
    tpooled = orc.roi_mean_pool([T(g[f"tfeat{l}"]) for l in range(4)], trois).view(B, Fm, -1)
    np.testing.assert_array_equal(tpooled.numpy(), g["tmpl_pooled"])
    pm, sc = T(g["prop_mask"]), T(g["prop_score"])
    out, loss, last = orc.dmm_container_forward(
        cfg, is_test, list(pooled.split(n_prop, 0)), [pm[b, :n_prop[b]] for b in range(B)],
        [sc[b, :n_prop[b]] for b in range(B)], list(tpooled.unbind(0)), T(g["tmpl_mask"]), T(g["valid"]),
        None if is_test else T(g["targets"]), expand=True, extra_frame=[int(v) for v in g["extra_frame"]])
    np.testing.assert_array_equal(out.numpy(), g["output_mask"])
    np.testing.assert_array_equal(last.numpy(), g["out_mask_last"])
    if not is_test:
        np.testing.assert_array_equal(np.array([float(x) for x in loss], np.float32), g["match_loss"])


@pytest.mark.parametrize("name", golden_names("hun_"))
def test_hungarian_layer_matches_reference(name):
    """algo='hun' (relax_match.py:120-126 + the head of match_model.py:125-147)."""
    g = load_golden(name)
    P, O, H, W, D, is_test = [int(v) for v in g["meta"]]
    cfg = default_cfg(20, 5, algo="hun")
    full, ms, ds, _, _ = orc.match_layer_forward(cfg, is_test, T(g["prop_feat"]), T(g["prop_mask"]), [T(g["tmpl_feat"])],
                                                 T(g["tmpl_mask"]), T(g["prop_score"]))
    np.testing.assert_array_equal(full.numpy(), g["full_outmask"])
    np.testing.assert_array_equal(ms.numpy(), g["match_score"])
    np.testing.assert_array_equal(ds.numpy(), g["det_score"])


def big_inputs(g):
    """Regenerates the headline-size inputs of a big_* golden and proves they are the bytes the reference saw."""
    import hashlib
    from dmm_net_b200.synth import make_problem
    P, O, H, W, D, mi, pi, is_test, config, index, lattice = [int(v) for v in g["meta"]]
    pr = make_problem(P, O, H, W, D, config=config, index=index)
    h = hashlib.sha256()
    for t in (pr.prop_feat, pr.prop_mask, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score):
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    assert h.hexdigest() == str(g["input_sha"]), "seeded generator no longer reproduces the golden's inputs"
    return pr, default_cfg(mi, pi), lattice


@pytest.mark.parametrize("name", golden_names("big_"))
def test_headline_size_layer_matches_reference(name):
    """50 x 10 x 256x448 (and 255x448): the oracle against the reference's outputs at the size the bench runs."""
    g = load_golden(name)
    pr, cfg, lattice = big_inputs(g)
    torch.set_num_threads(8)
    try:
        sim, _ = orc.cost_matrix(pr.prop_feat, pr.prop_mask, [pr.tmpl_feat], pr.tmpl_mask, cfg["score_weight"], None, expand=False)
        full, ms, ds, logic, bmat, R = orc.assign_and_apply(sim, pr.prop_mask, pr.prop_score, cfg["relax_max_iter"],
                                                            cfg["relax_proj_iter"], cfg["relax_learning_rate"], 1)
    finally:
        torch.set_num_threads(1)
    np.testing.assert_allclose(sim.numpy(), g["sim"], rtol=0, atol=1e-6)      # multi-threaded sums: not bit-pinned
    np.testing.assert_array_equal(logic.numpy(), g["logic"])
    np.testing.assert_allclose(bmat.numpy(), g["bmat"], rtol=0, atol=1e-5)
    flat = full.reshape(full.shape[0], -1)
    np.testing.assert_allclose(flat[:, ::lattice].numpy(), g["full_lattice"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(flat.double().sum(1).numpy(), g["full_rowsum"], rtol=1e-6)
