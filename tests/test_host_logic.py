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


def test_rows_around_the_layer_refuse_cpu_tensors_and_keep_the_reference_asserts():
    """K6-K9 host wrappers: no CPU fallback (RuntimeError before any kernel), shape asserts like the reference's"""
    from dmm_net_b200.utils.boxlist import BoxList
    from dmm_net_b200.utils.boxlist_ops import filter_results
    from dmm_net_b200.utils.masker import Masker
    m = torch.rand(2, 3, 8, 8)
    with pytest.raises(RuntimeError):
        ops.mask_pyramid(m, m, m, 4)
    with pytest.raises(RuntimeError):
        ops.merge_labels(m.view(2, 3, 64))
    with pytest.raises(RuntimeError):
        ops.paste_masks(torch.rand(3, 1, 28, 28), torch.rand(3, 4), 16, 16)
    with pytest.raises(RuntimeError):
        ops.box_nms(torch.rand(5, 4), torch.rand(5), 0.5)
    with pytest.raises(AssertionError):                          # init_pred_inst is [B,O,H,W] (trainer.py:184 CHECK4D)
        ops.mask_pyramid(m, m, m.view(2, 3, 64), 4)
    boxes = BoxList(torch.rand(3, 4), (16, 16))
    with pytest.raises(AssertionError):                          # masker.py:219 "Number of objects should be the same."
        Masker(0.5)([torch.rand(2, 1, 28, 28)], [boxes])
    with pytest.raises(AssertionError):                          # masker.py:214 "Masks and boxes should have the same length."
        Masker(0.5)([torch.rand(3, 1, 28, 28)] * 2, [boxes])
    assert filter_results([]) == []
    empty = BoxList(torch.zeros(0, 4), (16, 16))
    empty.add_field("scores", torch.zeros(0))
    assert len(filter_results([empty])[0]) == 0                  # nothing to suppress: no kernel, no sync


def test_boxlist_indexing_carries_fields():
    from dmm_net_b200.utils.boxlist import BoxList
    b = BoxList(torch.arange(20.).view(5, 4), (10, 10))
    b.add_field("scores", torch.arange(5.))
    b.add_field("mask", torch.arange(5)[:, None, None, None].expand(5, 1, 2, 2))
    k = b[torch.tensor([3, 0])]
    assert len(k) == 2 and k.size == (10, 10) and k.get_field("scores").tolist() == [3.0, 0.0]
    assert k.get_field("mask")[:, 0, 0, 0].tolist() == [3, 0] and k.bbox[0].tolist() == [12.0, 13.0, 14.0, 15.0]


def test_cosine_impl_names_and_packed_word_count():
    assert ops.COSINE_IMPLS == {"auto": 0, "simt": 1, "tc": 2}
    lib = __import__("dmm_net_b200._lib", fromlist=["load"]).load()
    assert lib.dmm_packed_words(256 * 448) == 3584
    assert lib.dmm_paste_masks_workspace_bytes(50) >= 50 * 16
    import ctypes
    h, w = ctypes.c_int(), ctypes.c_int()
    assert lib.dmm_mask_pyramid_level_size(255, 447, 3, ctypes.byref(h), ctypes.byref(w)) == 0 and (h.value, w.value) == (8, 14)
    assert lib.dmm_mask_pyramid_level_size(255, 447, 5, ctypes.byref(h), ctypes.byref(w)) == 1      # more than 5 levels


def test_hostmem_helpers_degrade_gracefully():
    """dmm_net_b200/hostmem.py: NUMA placement of pinned buffers is best effort -- on a single-node host (the GPU boxes are
    single-node VMs) everything must still work and say so."""
    from dmm_net_b200 import hostmem
    nodes = hostmem.memory_nodes()
    assert isinstance(nodes, list) and len(nodes) >= 1 and all(isinstance(n, int) for n in nodes)
    with hostmem.prefer_node(None) as ok:
        assert ok is False
    with hostmem.prefer_node(10 ** 6) as ok:                       # a node that does not exist: refused, not raised
        assert ok is False
    with hostmem.prefer_node(nodes[0]) as ok:
        t = torch.zeros(1024)
        assert isinstance(ok, bool) and float(t.sum()) == 0.0


def test_solver_shape_limits_are_reported_with_the_reference_flag_names():
    from dmm_net_b200 import ops
    lim = ops.limits()
    assert lim["max_templates"] == 16 and lim["max_solver_cols"] == 128
    ops.check_solver_shape(50, 10)
    with pytest.raises(RuntimeError, match="maxseqlen"):
        ops.check_solver_shape(50, 17)
    with pytest.raises(RuntimeError, match="sort_max_num"):
        ops.check_solver_shape(129, 5)
