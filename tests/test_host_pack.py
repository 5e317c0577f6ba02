"""CPU-side test of the host mask packer (runs without a GPU): bit i of word j = pixel 32j+i > 0.5, NaN -> 0, zero tail."""
import numpy as np
import pytest
import torch

from dmm_net_b200 import ops


@pytest.mark.parametrize("shape,md", [((3, 5, 7, 9), 2), ((2, 4, 1000), 1), ((1, 1, 31), 1), ((2, 3, 64), 1), ((6, 40, 56), 2),
                                      ((2, 0, 8, 8), 2)])
def test_host_pack_matches_numpy(shape, md):
    g = torch.Generator().manual_seed(sum(shape))
    m = torch.rand(shape, generator=g)
    if m.numel():
        m.view(-1)[0] = float("nan")
        m.view(-1)[-1] = 0.5                                     # exactly at the threshold: not set
    bits = ops.pack_masks_host(m, mask_dims=md)
    HW = int(np.prod(shape[-md:]))
    flat = m.reshape(-1, HW).numpy()
    want = np.packbits(flat > 0.5, axis=1, bitorder="little")
    want = np.pad(want, ((0, 0), (0, (-want.shape[1]) % 4))).view(np.uint32)
    got = bits.reshape(flat.shape[0], ops.packed_words(HW)).numpy().view(np.uint32)
    assert got.shape == want.shape and (got == want).all()
    assert bits.shape == tuple(shape[:-md]) + (ops.packed_words(HW),)
    for threads in (1, 3):
        assert torch.equal(ops.pack_masks_host(m, mask_dims=md, threads=threads), bits)


def test_host_pack_two_row_sets_one_team():
    """dmm_host_pack_masks2: proposal and template rows of a batch packed by one thread team == two separate calls"""
    import ctypes
    from dmm_net_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    HW = 37 * 53
    a, b = torch.rand(7 * 50, HW, generator=g), torch.rand(7 * 10, HW, generator=g)
    words = ops.packed_words(HW)
    da, db = torch.zeros(a.shape[0], words, dtype=torch.int32), torch.zeros(b.shape[0], words, dtype=torch.int32)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    for threads in (1, 5, 0):
        da.zero_(); db.zero_()
        assert lib.dmm_host_pack_masks2(vp(a), a.shape[0], vp(da), vp(b), b.shape[0], vp(db), HW, threads) == 0
        assert torch.equal(da, ops.pack_masks_host(a, mask_dims=1)) and torch.equal(db, ops.pack_masks_host(b, mask_dims=1))
    assert lib.dmm_host_pack_masks2(vp(a), a.shape[0], vp(da), None, 0, None, HW, 2) == 0      # an empty second set
    assert lib.dmm_host_pack_masks2(None, 3, None, None, 0, None, HW, 2) == 1                  # rows without a pointer
