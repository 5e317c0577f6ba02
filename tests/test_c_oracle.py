"""The plain-C integer oracle (oracle/iou_oracle.c) against the torch-CPU oracle, the reference's golden vectors, the host
packer and -- on the GPU -- K1's counts.  Three independent implementations of the same integers must agree exactly."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from dmm_net_b200 import ops
from dmm_net_b200.synth import make_problem
from oracle import match_oracle as orc


def c_oracle():
    import __graft_entry__ as ge
    lib = ctypes.CDLL(ge.build_c_oracle())
    lib.iou_oracle_pairwise.restype = ctypes.c_int
    lib.iou_oracle_pack_bits.restype = ctypes.c_int
    return lib


def c_pairwise(lib, prop, tmpl):
    P, O, HW = prop.shape[0], tmpl.shape[0], prop[0].numel()
    prop = np.ascontiguousarray(prop.reshape(P, HW).numpy(), np.float32)
    tmpl = np.ascontiguousarray(tmpl.reshape(O, HW).numpy(), np.float32)
    inter, at, apr = np.zeros((O, P), np.int64), np.zeros(O, np.int64), np.zeros(P, np.int64)
    iou = np.zeros((O, P), np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.iou_oracle_pairwise(vp(prop), vp(tmpl), P, O, ctypes.c_longlong(HW), vp(inter), vp(at), vp(apr), vp(iou))
    assert rc == 0
    return inter, at, apr, iou


@pytest.mark.parametrize("name", golden_names("layer_"))
def test_c_oracle_reproduces_reference_iou(name):
    g = load_golden(name)
    lib = c_oracle()
    _, _, _, iou = c_pairwise(lib, torch.from_numpy(g["prop_mask"]), torch.from_numpy(g["tmpl_mask"]))
    assert np.array_equal(iou, g["iou"])                                    # bit-equal to the unmodified reference


def test_c_oracle_equals_torch_oracle_and_host_packer():
    lib = c_oracle()
    pr = make_problem(13, 4, 37, 53, 8, config=1, index=3)
    pr.prop_mask[1].zero_()
    pr.prop_mask[2].fill_(0.5)                                              # at the threshold: not set
    inter, at, apr, iou = c_pairwise(lib, pr.prop_mask, pr.tmpl_mask)
    want = orc.pairwise_binary_iou(pr.prop_mask.view(13, -1), pr.tmpl_mask.view(4, -1), expand=False)
    assert np.array_equal(iou, want.numpy())
    assert apr[1] == 0 and apr[2] == 0
    HW = 37 * 53
    rows = np.ascontiguousarray(pr.prop_mask.view(13, HW).numpy())
    bits = np.zeros((13, ops.packed_words(HW)), np.uint32)
    lib.iou_oracle_pack_bits(rows.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(13), ctypes.c_longlong(HW),
                             bits.ctypes.data_as(ctypes.c_void_p))
    got = ops.pack_masks_host(pr.prop_mask.view(13, HW), mask_dims=1).numpy().view(np.uint32)
    assert np.array_equal(bits, got)
    assert np.array_equal(np.array([bin(int(w)).count("1") for w in bits.reshape(-1)]).reshape(13, -1).sum(1), apr)


@pytest.mark.gpu
def test_k1_counts_equal_c_oracle():
    lib = c_oracle()
    pr = make_problem(50, 10, 64, 112, 8, config=2, index=5)
    inter, at, apr, iou = c_pairwise(lib, pr.prop_mask, pr.tmpl_mask)
    r = ops.mask_iou_pairwise(pr.prop_mask[None].cuda(), pr.tmpl_mask[None].cuda(), want_counts=True)
    counts = r["counts"][0].cpu().numpy()
    assert np.array_equal(counts[:500].reshape(10, 50), inter)
    assert np.array_equal(counts[500:510], at) and np.array_equal(counts[510:], apr)
    assert np.array_equal(r["iou"][0].cpu().numpy(), iou)
