"""GPU parity: CSR/COO/MTX -> bmSparse through the C ABI vs the oracle (bit-exact structure and values)."""
import os

import numpy as np
import pytest
import torch

from tests.util import assert_structure_equal, csr_rows, hex_to_u64, load_golden, random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import bmsparse_spgemm_spmv_b200 as B
    return B


def _check(B, O, nr, nc, rp, ci, v, transposed, dtype=torch.float16, device_input=False):
    if device_input:
        M = B.bmSpMatrix.from_csr(nr, nc, torch.tensor(rp, device="cuda"), torch.tensor(ci, device="cuda"),
                                  torch.tensor(v, device="cuda"), transpose=transposed, dtype=dtype)
    else:
        M = B.bmSpMatrix.from_csr(nr, nc, rp, ci, v, transpose=transposed, dtype=dtype)
    exp = O.csr_to_bmsp(nr, nc, rp, ci, v, transposed=transposed, f16=(dtype == torch.float16))
    k, b, o, vals = M.download()
    assert_structure_equal((k, b, o), exp, f"{nr}x{nc} t={transposed}")
    assert np.array_equal(vals.astype(np.float32), exp.values), "values differ"
    # derived compact arrays
    brp = M.block_row_ptr.cpu().numpy()
    nbr = (nr + 7) // 8
    exp_brp = np.searchsorted(exp.keys >> np.uint64(32), np.arange(nbr + 1, dtype=np.uint64), side="left")
    assert np.array_equal(brp, exp_brp)
    assert np.array_equal(M.block_col.cpu().numpy(), (exp.keys & np.uint64(0xFFFFFFFF)).astype(np.int32))
    return M, exp


def test_fixture(B, oracle):
    g = load_golden("ragusa16.json")
    for name, tr in (("A", False), ("B", True)):
        M = B.bmSpMatrix.from_coo(g["num_rows"], g["num_cols"], g[name]["rows"], g[name]["cols"], g[name]["vals"], transpose=tr)
        k, b, o, v = M.download()
        pre = "A" if name == "A" else "Bt"
        assert np.array_equal(k, hex_to_u64(g[f"{pre}_keys"]))
        assert np.array_equal(b, hex_to_u64(g[f"{pre}_bmps"]))
        assert o.tolist() == g[f"{pre}_offsets"]
        assert M.num_rows == 24 and M.nnz == 81 and M.block_num == 9


@pytest.mark.parametrize("shape", [(24, 24), (61, 29), (100, 7), (7, 300), (257, 513), (1000, 1000)])
@pytest.mark.parametrize("transposed", [False, True])
def test_random(B, oracle, shape, transposed):
    nr, nc = shape
    rp, ci, v = random_csr(nr, nc, 0.08, seed=nr * 31 + nc, empty_block_rows=(1,) if nr > 24 else ())
    _check(B, oracle, nr, nc, rp, ci, v, transposed)
    _check(B, oracle, nr, nc, rp, ci, v, transposed, device_input=True)


def test_fp32_values_and_dense_blocks(B, oracle):
    rp, ci, v = random_csr(200, 200, 0.6, seed=5, fp16=False)
    _check(B, oracle, 200, 200, rp, ci, v, False, dtype=torch.float32)
    _check(B, oracle, 200, 200, rp, ci, v, True, dtype=torch.float16)


def test_poisson_and_long_rows(B, oracle):
    nr, nc, rp, ci, v = B.generators.poisson5pt(64, 48)
    _check(B, oracle, nr, nc, rp, ci, v, False)
    _check(B, oracle, nr, nc, rp, ci, v, True)
    # one very long row and empty rows around it
    rng = np.random.default_rng(3)
    cols = np.sort(rng.choice(20000, 6000, replace=False)).astype(np.int32)
    rp = np.zeros(41, np.int32); rp[18:] = 6000
    rp2 = rp.copy(); extra = np.array([5, 9, 19999], np.int32)
    ci = np.concatenate([cols, extra]); rp2[31:] += 3
    v = np.arange(1, ci.size + 1, dtype=np.float32) % 7 + 1
    _check(B, oracle, 40, 20000, rp2, ci, v, False)
    _check(B, oracle, 40, 20000, rp2, ci, v, True)


def test_empty_and_tiny(B, oracle):
    M = B.bmSpMatrix.from_csr(16, 16, np.zeros(17, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32))
    assert M.block_num == 0 and M.nnz == 0
    rp = np.array([0, 1], np.int32); ci = np.array([0], np.int32); v = np.array([2.0], np.float32)
    M, exp = _check(B, oracle, 1, 1, rp, ci, v, False)
    assert M.download()[1][0] == np.uint64(1 << 63)


def test_rejects_bad_input(B):
    rp = np.array([0, 2], np.int32)
    with pytest.raises(B.BmspError) as e:
        B.bmSpMatrix.from_csr(1, 8, rp, np.array([3, 3], np.int32), np.ones(2, np.float32))
    assert e.value.code == 4       # duplicate
    with pytest.raises(B.BmspError) as e:
        B.bmSpMatrix.from_csr(1, 8, rp, np.array([5, 3], np.int32), np.ones(2, np.float32))
    assert e.value.code == 3       # unsorted
    with pytest.raises(B.BmspError) as e:
        B.bmSpMatrix.from_csr(1, 8, rp, np.array([5, 8], np.int32), np.ones(2, np.float32))
    assert e.value.code == 8       # range


def test_mtx_symmetric_and_roundtrip(B, oracle, tmp_path):
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% c\n5 5 4\n1 1 2.5\n3 1 -1\n5 4 7\n5 5 1e-3\n")
    M = B.bmSpMatrix(str(p), False)
    nr, nc, r, c, v = oracle.read_mtx(str(p))
    exp = oracle.coo_to_bmsp(nr, nc, r, c, v)
    k, b, o, vals = M.download()
    assert_structure_equal((k, b, o), exp)
    assert M.nnz == 6
    rr, cc, vv = M.generate_coo()
    er, ec, ev = oracle.bmsp_to_coo(exp)
    assert np.array_equal(rr, er) and np.array_equal(cc, ec) and np.array_equal(vv, ev)
    assert M.compare(er, ec, ev) == (0, 0, 0.0, 0.0)


def test_from_arrays_and_block_transpose(B, oracle):
    rp, ci, v = random_csr(150, 90, 0.15, seed=9)
    A = oracle.csr_to_bmsp(150, 90, rp, ci, v, transposed=False)
    At = oracle.csr_to_bmsp(150, 90, rp, ci, v, transposed=True)
    M = B.bmSpMatrix.from_arrays(150, 90, A.block_num, A.keys, A.bmps, A.offsets, A.values.astype(np.float16))
    k, b, o, vals = M.download()
    assert_structure_equal((k, b, o), A)
    T = M.block_transpose()
    assert T.transposed
    k, b, o, vals = T.download()
    assert_structure_equal((k, b, o), At)
    assert np.array_equal(vals.astype(np.float32), At.values)
    k2, b2, o2, v2 = T.block_transpose().download()
    assert_structure_equal((k2, b2, o2), A)
    assert np.array_equal(v2.astype(np.float32), A.values)
    # slicing block rows
    S = M.slice_block_rows(3, 11)
    ks, bs, os_, vs = S.download()
    sel = (A.keys >> np.uint64(32) >= 3) & (A.keys >> np.uint64(32) < 11)
    assert np.array_equal(ks, A.keys[sel]) and np.array_equal(bs, A.bmps[sel])
    assert np.array_equal(os_, A.offsets[sel] - A.offsets[sel][0])


def test_device_csr_row_ptr_is_validated(B):
    """a device CSR with a bad row_ptr must be rejected, not binary-searched out of bounds."""
    import bmsparse_spgemm_spmv_b200._lib as L
    nr, nc = 64, 64
    ci = torch.arange(64, dtype=torch.int32, device="cuda"); v = torch.ones(64, device="cuda")
    good = torch.arange(65, dtype=torch.int32, device="cuda")
    assert B.bmSpMatrix.from_csr(nr, nc, good, ci, v).nnz == 64
    swapped = good.clone(); swapped[10], swapped[11] = good[11].item(), good[10].item()      # decreases once
    short = good.clone(); short[-1] = 63                                                       # does not end at nnz
    for bad in (good + 1, swapped, short, -good):
        with pytest.raises(L.BmspError) as e:
            B.bmSpMatrix.from_csr(nr, nc, bad, ci, v)
        assert e.value.code == 1
