"""GPU parity against the goldens of oracle/make_golden_container.py (all produced by the UNMODIFIED reference):
the batched DMM_Model container vs the reference's per-video loop (dmm_model.py:22-158), algo='hun'
(relax_match.py:120-126), and the layer at the headline size (50 x 10 x 256x448 / 255x448)."""
import numpy as np
import pytest
import torch

from conftest import T, golden_names, load_golden
from dmm_net_b200 import ops
from dmm_net_b200.modules.dmm_model import DMM_Model
from dmm_net_b200.modules.match_model import MatchModel
from dmm_net_b200.synth import default_cfg
from dmm_net_b200.utils.boxlist import BoxList
from test_oracle_golden import big_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


def close(a, b, tol, what):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    err = float(np.abs(a - np.asarray(b)).max()) if a.size else 0.0
    assert a.shape == np.asarray(b).shape, (what, a.shape, np.asarray(b).shape)
    assert err <= tol, f"{what}: max abs err {err:.3e} > {tol}"


@pytest.mark.parametrize("name", golden_names("container_"))
def test_container_against_reference_golden(name):
    """backbone features + boxes in -> (K5, K2, K1 through the pointer table, K3, K4 with the fused row scatter) ->
    output_mask / out_mask_last / match_loss (and gradients into the feature maps) of the reference container."""
    g = load_golden(name)
    B, P, Fm, H, W, C, mi, pi, is_test = [int(v) for v in g["meta"]]
    n_prop = [int(v) for v in g["n_prop"]]
    cfg = default_cfg(mi, pi)
    model = DMM_Model(cfg, is_test=is_test).to(DEV)
    feats = tuple(T(g[f"feat{l}"], DEV).requires_grad_(not is_test) for l in range(4))
    tfeats = tuple(T(g[f"tfeat{l}"], DEV) for l in range(4))
    valid = T(g["valid"], DEV)
    props = []
    for b in range(B):
        bl = BoxList(T(g[f"boxes{b}"]), (W, H))
        bl.add_field("mask", T(g["prop_mask"])[b, :n_prop[b]].unsqueeze(1))
        bl.add_field("scores", T(g["prop_score"])[b, :n_prop[b]])
        props.append(bl.to(DEV))
    tplt = model.fill_template_dict(None, [BoxList(T(g[f"tboxes{b}"]), (W, H)).to(DEV) for b in range(B)],
                                    {"backbone_feature": tfeats, "refine_input_feat": tfeats}, None, valid)
    close(torch.stack([tplt[b]["feat"][0] for b in range(B)], 0), g["tmpl_pooled"], 1e-5, "template features (K5)")
    close(model.feature_extractor(tuple(f.detach() for f in feats), props), g["prop_pooled"], 1e-5, "proposal features (K5)")
    mask_last = T(g["tmpl_mask"], DEV)
    if is_test:
        with torch.no_grad():
            out, _, loss, last = model.inference({"args": None, "shape": (H, W), "valid": valid,
                                                  "extra_frame": [int(v) for v in g["extra_frame"]]}, props, feats, mask_last, tplt)
        assert loss == []
    else:
        out, _, loss, last = model(None, props, feats, mask_last, tplt, valid, T(g["targets"], DEV))
        assert len(loss) == B
        close(torch.stack(loss), g["match_loss"], 1e-5, "match_loss")
    close(out, g["output_mask"], TOL, "output_mask")
    close(last, g["out_mask_last"], TOL, "out_mask_last")
    invalid_rows = (np.arange(Fm)[None, :] >= g["valid"].sum(1)[:, None]) | (g["valid"] == 0)
    assert float(out.detach().cpu()[torch.from_numpy(invalid_rows)].abs().sum()) == 0.0
    if not is_test:
        ((out * T(g["w_mask"], DEV)).sum() + 3.0 * sum(loss)).backward()
        for l in range(4):
            want = g[f"g_feat{l}"]
            close(feats[l].grad, want, TOL * max(1.0, float(np.abs(want).max())), f"d/d feature level {l}")


@pytest.mark.parametrize("name", golden_names("hun_"))
def test_hungarian_algo_against_reference_golden(name):
    g = load_golden(name)
    P, O, H, W, D, is_test = [int(v) for v in g["meta"]]
    layer = MatchModel(default_cfg(20, 5, algo="hun"), is_test=is_test)
    pf, tf, sc = T(g["prop_feat"], DEV), T(g["tmpl_feat"], DEV), T(g["prop_score"], DEV)
    pm, tm = T(g["prop_mask"], DEV), T(g["tmpl_mask"], DEV)
    with torch.no_grad():
        sim, n_prop, n_tplt, _ = layer.compute_cost_matrix({"proposed": pf, "template": [tf]}, {"proposed": pm, "template": tm},
                                                           {"proposal_score": sc}, None)
        close(sim, g["sim"], 5e-6, "sim")
        _, _, _, logic, bmat = layer.match_with_first_frame(sim, P, O, pm, sc, tm)
        full, ms, ds, full2, loss = layer(pf, pm, [tf], tm, sc)
    assert full is full2 and loss == {}
    np.testing.assert_array_equal(logic.cpu().numpy(), g["logic"])
    np.testing.assert_array_equal(bmat.cpu().numpy(), g["bmat"])                     # one-hot: exact
    close(full, g["full_outmask"], 1e-6, "full_outmask")
    close(ms, g["match_score"], 5e-6, "match_score")
    close(ds, g["det_score"], 1e-6, "det_score")


@pytest.mark.parametrize("name", golden_names("big_"))
def test_headline_size_against_reference_golden(name):
    """BASELINE configs[1] exactly (and the scripts' 255x448 with eval.yaml's 40x5) against the reference itself."""
    g = load_golden(name)
    pr, cfg, lattice = big_inputs(g)                                                 # asserts the inputs' sha256
    pr = pr.to(DEV)
    layer = MatchModel(cfg, is_test=1)
    with torch.no_grad():
        out = layer.forward_many(pr.prop_feat[None], pr.prop_mask[None], pr.tmpl_feat[None], pr.tmpl_mask[None],
                                 pr.prop_score[None])
    assert int(out["n_list"][0]) == int(g["n_list"])                                 # shipped presets: no exit slack
    close(out["sim"][0], g["sim"], 5e-6, "sim")
    np.testing.assert_array_equal(out["logic"][0].cpu().numpy(), g["logic"])
    close(out["Bmat"][0], g["bmat"], TOL, "bmat")
    assert np.array_equal(out["Bmat"][0].argmax(1).cpu().numpy(), g["bmat"].argmax(1))
    close(out["match_score"][0], g["match_score"], TOL, "match_score")
    close(out["det_score"][0], g["det_score"], TOL, "det_score")
    flat = out["full_outmask"][0].reshape(out["full_outmask"].shape[1], -1)
    close(flat[:, ::lattice], g["full_lattice"], TOL, "full_outmask (lattice)")
    np.testing.assert_allclose(flat.double().sum(1).cpu().numpy(), g["full_rowsum"], rtol=1e-5)
