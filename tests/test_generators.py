"""CPU suite: the torch evaluation of the counter-based generators (used on the GPU for the big configs) is bit-identical
to the numpy one the oracle tests use."""
import numpy as np

from bmsparse_spgemm_spmv_b200 import generators as G


def test_values_fp16_torch_matches_numpy():
    a = G.values_fp16(4, 200000)
    b = G.values_fp16_torch(4, 200000, "cpu").numpy()
    assert np.array_equal(a, b)


def test_rmat_torch_matches_numpy():
    for scale in (6, 11):
        n, _, rp, ci, v = G.rmat(scale)
        n2, _, rp2, ci2, v2 = G.rmat_torch(scale, device="cpu")
        assert n == n2
        assert np.array_equal(rp, rp2.numpy()) and np.array_equal(ci, ci2.numpy()) and np.array_equal(v, v2.numpy())
