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
