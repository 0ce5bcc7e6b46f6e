"""CPU: the numpy restatement of libmobgt's attention-dropout keep mask (tests/helpers.attn_drop_keep, mirroring
mobgt_b200/csrc/common.cuh attn_drop_*) behaves like nn.Dropout's Bernoulli mask (model_fqandtoyo.py:1674, 1704): keep rate
1 - p, no correlation between neighbouring elements / rows / planes / seeds, deterministic in (seed, plane, row, col).
The GPU parity test (tests/test_k2_k3_k4.py) checks that the kernels draw exactly this mask."""
import numpy as np

from helpers import attn_drop_keep


def _corr(a, b):
    a, b = a - a.mean(), b - b.mean()
    return float((a * b).mean() / np.sqrt((a * a).mean() * (b * b).mean()))


def test_keep_rate_scale_and_determinism():
    seed = 0x9E3779B97F4A7C15
    K = np.stack([attn_drop_keep(pl, 129, 0.1, seed)[0] for pl in range(48)]).astype(np.float64)
    scale = attn_drop_keep(0, 129, 0.1, seed)[1]
    assert K.shape == (48, 129, 129)
    assert abs(K.mean() - 0.9) < 3e-3
    assert abs(scale - 1.0 / 0.9) < 2e-4                       # 65536 / (65536 - round(0.1 * 65536))
    assert abs(K.mean() * scale - 1.0) < 4e-3                  # unbiased in expectation
    assert np.array_equal(attn_drop_keep(7, 129, 0.1, seed)[0], K[7].astype(bool))
    assert not np.array_equal(attn_drop_keep(7, 129, 0.1, seed + 1)[0], K[7].astype(bool))
    # a smaller graph sees the same mask on its own rows / columns (the mask depends on (row, col), not on T)
    assert np.array_equal(attn_drop_keep(7, 40, 0.1, seed)[0], K[7, :40, :40].astype(bool))


def test_no_structure_between_neighbours():
    seed = 0x0123456789ABCDEF
    K = np.stack([attn_drop_keep(pl, 136, 0.25, seed)[0] for pl in range(64)]).astype(np.float64)
    assert abs(K.mean() - 0.75) < 3e-3
    assert abs(_corr(K[:, :, :-1], K[:, :, 1:])) < 6e-3        # adjacent columns
    assert abs(_corr(K[:, :-1], K[:, 1:])) < 6e-3              # adjacent rows
    assert abs(_corr(K[:-1], K[1:])) < 6e-3                    # adjacent planes
    for d in range(1, 8):                                      # the 8 fields derived from one hash word group
        assert abs(_corr(K[:, :, 0::8], K[:, :, d::8])) < 1.2e-2
    per_col = K.mean((0, 1))
    assert per_col.min() > 0.72 and per_col.max() < 0.78
    K2 = np.stack([attn_drop_keep(pl, 136, 0.25, seed + 1)[0] for pl in range(64)]).astype(np.float64)
    assert abs(_corr(K, K2)) < 6e-3                            # consecutive call seeds


def test_p_zero_keeps_everything():
    keep, scale = attn_drop_keep(3, 50, 0.0, 123)
    assert keep.all() and scale == 1.0
