"""CPU-side host logic: shape helpers, the reference's AssertionError convention, config handling."""
import pytest
import torch

from dmm_net_b200 import ops
from dmm_net_b200.modules.match_model import MatchModel
from dmm_net_b200.synth import default_cfg, make_problem
from dmm_net_b200.utils import checker


def test_pad_rule_and_packed_words():
    assert ops.pad_cols(50, 10) == 50 and ops.pad_cols(3, 4) == 5 and ops.pad_cols(4, 4) == 5 and ops.pad_cols(5, 4) == 5
    assert [ops.packed_words(n) for n in (0, 1, 32, 33, 114688)] == [0, 1, 1, 2, 3584]
    assert ops.host_threads() >= 1


def test_checker_raises_assertion_error_like_the_reference():
    t = torch.zeros(2, 3)
    assert tuple(checker.CHECK2D(t)) == (2, 3)
    for fn in (checker.CHECK3D, checker.CHECK4D, checker.CHECK5D):
        with pytest.raises(AssertionError):
            fn(t)
    checker.CHECKEQ(3, 3)
    with pytest.raises(AssertionError):
        checker.CHECKEQ(3, 4)
    with pytest.raises(AssertionError):
        checker.CHECKSIZE(t, (3, 2))


def test_layer_reads_the_five_config_keys_and_has_no_state():
    cfg = default_cfg(40, 5, 0.2, 0.25)
    layer = MatchModel(cfg, is_test=1)
    assert (layer.max_iter, layer.proj_iter, layer.relax_lr, layer.match_algo, layer.is_test) == (40, 5, 0.2, "relax", 1)
    assert len(layer.state_dict()) == 0 and len(list(layer.parameters())) == 0      # checkpoints are unaffected
    with pytest.raises(AssertionError):
        MatchModel(default_cfg(algo="sinkhorn"))
    with pytest.raises(KeyError):
        MatchModel({"matching": {"algo": "relax"}})


def test_shape_asserts_fire_before_any_kernel():
    pr = make_problem(4, 2, 8, 8, 16)
    layer = MatchModel(default_cfg(), 1)
    with pytest.raises(AssertionError):                          # proposed_mask must be [P,H,W]
        layer(pr.prop_feat, pr.prop_mask[0], [pr.tmpl_feat], pr.tmpl_mask, pr.prop_score)
    with pytest.raises(AssertionError):                          # one score per proposal
        layer(pr.prop_feat, pr.prop_mask, [pr.tmpl_feat], pr.tmpl_mask, pr.prop_score[:3])
