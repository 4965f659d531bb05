"""CPU: the product's synthetic workload generators equal the oracle's (bench inputs == test inputs)."""
import numpy as np
import torch

import codon_oracle as orc
from codon_b200 import synthetic as syn


def test_generators_agree():
    for scale, seed in ((4, 0), (8, 2), (16, 1)):
        a, b = orc.synthetic_state_dict(scale, seed), syn.synthetic_state_dict(scale, seed)
        assert a.keys() == b.keys()
        assert all(torch.equal(a[k], b[k]) for k in a)
    xa, ya = orc.synthetic_frames(2, 19, 23, 5)
    xb, yb = syn.synthetic_frames(2, 19, 23, 5)
    assert torch.equal(xa, xb) and torch.equal(ya, yb)
    assert syn.FLOPS_PER_PIXEL == orc.flops_per_pixel()
    assert float(xa.min()) >= 0 and float(xa.max()) <= 1
    np.testing.assert_array_equal(np.round(xa.numpy() * 255), xa.numpy() * 255)
