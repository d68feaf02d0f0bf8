"""CPU suite: pins the oracle to the reference's fixture, to the survey's independently derived vectors,
and to the reference's own cusp host kernels (golden JSON written by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from tests.util import hex_to_u64, load_golden


@pytest.fixture(scope="module")
def rag():
    return load_golden("ragusa16.json")


def _mats(O, g):
    A = O.coo_to_bmsp(g["num_rows"], g["num_cols"], g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])
    Bt = O.coo_to_bmsp(g["num_rows"], g["num_cols"], g["B"]["rows"], g["B"]["cols"], g["B"]["vals"], transposed=True)
    return A, Bt


def test_fixture_conversion_matches_golden_and_survey(oracle, rag):
    A, Bt = _mats(oracle, rag)
    assert np.array_equal(A.keys, hex_to_u64(rag["A_keys"]))
    assert np.array_equal(A.bmps, hex_to_u64(rag["A_bmps"]))
    assert np.array_equal(A.bmps, hex_to_u64(rag["survey_A_bmps"]))
    assert A.offsets.tolist() == rag["A_offsets"] == rag["survey_A_offsets"]
    assert np.array_equal(Bt.keys, hex_to_u64(rag["Bt_keys"]))
    assert np.array_equal(Bt.bmps, hex_to_u64(rag["survey_Bt_bmps"]))
    assert A.block_num == 9 and A.nnz == 81


def test_fixture_spmv(oracle, rag):
    A, _ = _mats(oracle, rag)
    y = oracle.spmv(A, np.ones(rag["num_cols"], np.float32))
    assert y.tolist() == rag["survey_spmv_ones"] == rag["spmv_ones"]


def test_fixture_spgemm(oracle, rag):
    A, Bt = _mats(oracle, rag)
    Cm = oracle.spgemm(A, Bt)
    assert np.array_equal(Cm.keys, hex_to_u64(rag["C_keys"]))
    assert np.array_equal(Cm.bmps, hex_to_u64(rag["survey_C_bmps"]))
    assert Cm.offsets.tolist() == rag["survey_C_offsets"] and Cm.nnz == 255
    r, c, v = oracle.bmsp_to_coo(Cm)
    got = {(int(a), int(b)): float(x) for a, b, x in zip(r, c, v)}
    exp = {(a, b): x for a, b, x in zip(rag["C_coo"]["rows"], rag["C_coo"]["cols"], rag["C_coo"]["vals"])}
    assert got == exp          # small integers: exact


def test_decode_roundtrip_both_orientations(oracle):
    rng = np.random.default_rng(0)
    for (nr, nc) in [(40, 40), (61, 29), (100, 7)]:
        mask = rng.random((nr, nc)) < 0.2
        r, c = np.nonzero(mask)
        v = rng.integers(1, 100, r.size).astype(np.float32)
        for tr in (False, True):
            M = oracle.coo_to_bmsp(nr, nc, r, c, v, transposed=tr)
            rr, cc, vv = oracle.bmsp_to_coo(M)
            o = np.lexsort((cc, rr))
            assert np.array_equal(rr[o], r) and np.array_equal(cc[o], c) and np.array_equal(vv[o], v)
            assert np.array_equal(M.offsets, np.concatenate([[0], np.cumsum([bin(int(b)).count("1") for b in M.bmps])[:-1]]).astype(np.uint64))


def test_pair_bitmap_against_definition(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = int(rng.integers(0, 2**63)) ^ (int(rng.integers(0, 2)) << 63)
        bt = int(rng.integers(0, 2**63)) & int(rng.integers(0, 2**63))
        exp = 0
        for i in range(8):
            ai = (a >> (56 - 8 * i)) & 0xFF
            for j in range(8):
                bj = (bt >> (56 - 8 * j)) & 0xFF
                if ai & bj:
                    exp |= 1 << (63 - (i * 8 + j))
        assert oracle.pair_bitmap(a, bt) == exp


def test_f16_rounding_matches_numpy(oracle):
    rng = np.random.default_rng(2)
    v = np.concatenate([rng.uniform(-70000, 70000, 5000), rng.uniform(-1e-4, 1e-4, 5000), [0.0, 1.0, 65504.0, 6e-8]]).astype(np.float32)
    with np.errstate(over="ignore"):
        assert np.array_equal(oracle.f16_round(v), v.astype(np.float16).astype(np.float32))


def test_cusp_port_matches_reference_kernels(oracle):
    g = load_golden("cusp_host.json")
    for name, c in g["cases"].items():
        N = c["m"] * c["n"]
        rp, ci, v = (np.array(c[k]) for k in ("rp", "ci", "v"))
        orp, oci, ov = oracle.poisson5pt(c["m"], c["n"])
        assert np.array_equal(orp, rp) and np.array_equal(oci, ci) and np.array_equal(ov, v), name
        y = oracle.csr_spmv(rp, ci, v, np.array(c["x"], np.float32))
        assert y.tolist() == c["y"], name
        for key, drop, thr in (("seq", True, 1), ("omp", False, 3)):
            crp, cci, cv = oracle.csr_spgemm(N, N, rp, ci, v, rp, ci, v, drop_zeros=drop, threads=thr)
            assert crp.tolist() == c[key]["rp"] and cci.tolist() == c[key]["ci"] and cv.tolist() == c[key]["v"], (name, key)


def test_cusp_reference_library_if_built(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rp, ci, v = oracle.poisson5pt(13, 9)
    N = 13 * 9
    a = oracle.csr_spgemm(N, N, rp, ci, v, rp, ci, v)
    b = oracle.ref_csr_spgemm(N, N, rp, ci, v, N, N, rp, ci, v)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))


def test_config1_poisson256_counts(oracle):
    """BASELINE config 1 sizes: nnz(A) = 326 656, nnz(A*A) = 846 852, 40 384 blocks."""
    from bmsparse_spgemm_spmv_b200 import generators as G
    nr, nc, rp, ci, v = G.poisson5pt(256, 256)
    orp, oci, ov = oracle.poisson5pt(256, 256)
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(v, ov)
    assert ci.size == 326656
    crp, cci, cv = oracle.csr_spgemm(nr, nc, rp, ci, v, rp, ci, v)
    assert cci.size == 846852
    A = oracle.csr_to_bmsp(nr, nc, rp, ci, v)
    assert A.block_num == 40384


def test_generators_shapes():
    from bmsparse_spgemm_spmv_b200 import generators as G
    n, _, rp, ci, v = G.uniform_random(1000, 16, seed=2)
    assert ci.size == 16000 and np.all(np.diff(ci.reshape(1000, 16), axis=1) > 0)
    n, _, rp, ci, v = G.block_clustered(64)
    assert n == 512 and rp[-1] == ci.size and ci.min() >= 0 and ci.max() < n
    for r in range(0, n, 37):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)
    n, _, rp, ci, v = G.rmat(10)
    assert n == 1024 and rp[-1] == ci.size
    for r in range(0, n, 41):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)
    assert np.all(v.astype(np.float16).astype(np.float32) == v)


def test_oracle_matches_recorded_reference_cuda_outputs(oracle):
    """tests/golden/ref_cuda_golden.json was written on the B200 box by tests/test_gpu_reference_cuda.py from the
    reference's OWN bmSparse_mult / bmSparse_SpMV (compiled for sm_100 into oracle/_ref).  The oracle must reproduce
    the recorded structure fingerprints exactly and the value sums within the reference's fp16-product rounding."""
    from tests.util import random_csr
    O = oracle
    rec = load_golden("ref_cuda_golden.json")["cases"]
    g = load_golden("ragusa16.json")
    mats = {"ragusa16_AxB": (O.coo_to_bmsp(24, 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"]),
                             O.coo_to_bmsp(24, 24, g["B"]["rows"], g["B"]["cols"], g["B"]["vals"], transposed=True))}
    rp, ci, v = O.poisson5pt(64, 64)
    mats["poisson64_AxA"] = (O.csr_to_bmsp(4096, 4096, rp, ci, v), O.csr_to_bmsp(4096, 4096, rp, ci, v, transposed=True))
    a32 = O.csr_to_bmsp(4096, 4096, rp, ci, v, f16=False)
    rp, ci, v = random_csr(256, 256, 0.06, seed=77)
    mats["random256_AxA"] = (O.csr_to_bmsp(256, 256, rp, ci, v), O.csr_to_bmsp(256, 256, rp, ci, v, transposed=True))
    for name, (a, bt) in mats.items():
        exp = O.spgemm(a, bt); r = rec[name]
        assert exp.block_num == r["C_blocks"] and exp.nnz == r["C_nnz"] == r["offsets_last"], name
        assert int(np.bitwise_xor.reduce(exp.keys)) == r["keys_crc"] and int(np.bitwise_xor.reduce(exp.bmps)) == r["bmps_crc"], name
        assert abs(float(exp.values.sum()) - r["values_sum"]) <= 2e-3 * float(np.abs(exp.values).sum()) + 1e-6, name
    y = O.spmv(a32, np.ones(4096, np.float32))
    assert float(y.sum()) == rec["poisson64_spmv_ones"]["y_sum"] and float(np.abs(y).sum()) == rec["poisson64_spmv_ones"]["y_abs_sum"]


def test_openmp_variants_match_the_scalar_oracle(oracle):
    """oracle/bmsp_oracle.c's block-row-parallel conversion / SpMV / SpGEMM (used for the full-size GPU parity tests) are the same
    functions as the scalar, reference-following ones: bit for bit on random matrices with empty block rows and on the fixture."""
    O = oracle
    from tests.util import random_csr
    cases = [(24, 24, 0.3, 1, ()), (61, 29, 0.2, 2, ()), (513, 300, 0.05, 3, (0, 2)), (1000, 1000, 0.01, 4, (3,)), (200, 2049, 0.3, 5, ())]
    for nr, nc, den, seed, emp in cases:
        rp, ci, v = random_csr(nr, nc, den, seed, emp)
        for tr in (False, True):
            a = O.csr_to_bmsp(nr, nc, rp, ci, v, transposed=tr); b = O.csr_to_bmsp_omp(nr, nc, rp, ci, v, transposed=tr)
            assert np.array_equal(a.keys, b.keys) and np.array_equal(a.bmps, b.bmps) and np.array_equal(a.offsets, b.offsets)
            assert np.array_equal(a.values, b.values)
        rp2, ci2, v2 = random_csr(nc, nr, den, seed + 10)
        A = O.csr_to_bmsp(nr, nc, rp, ci, v); Bt = O.csr_to_bmsp(nc, nr, rp2, ci2, v2, transposed=True)
        c1 = O.spgemm(A, Bt); c2 = O.spgemm_omp(A, Bt)
        assert np.array_equal(c1.keys, c2.keys) and np.array_equal(c1.bmps, c2.bmps) and np.array_equal(c1.offsets, c2.offsets)
        assert np.array_equal(c1.values.astype(np.float32), c2.values)
        x = np.random.default_rng(0).uniform(-1, 1, nc).astype(np.float32)
        assert np.array_equal(O.spmv(A, x), O.spmv_omp(A, x))
    g = load_golden("ragusa16.json")
    A = O.coo_to_bmsp(24, 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])
    Bt = O.coo_to_bmsp(24, 24, g["B"]["rows"], g["B"]["cols"], g["B"]["vals"], transposed=True)
    c = O.spgemm_omp(A, Bt)
    assert c.nnz == 255 and c.block_num == 9


def test_sha256_of_structure_arrays(oracle):
    import hashlib
    a = np.arange(1000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    assert oracle.sha256_u64(a) == hashlib.sha256(a.tobytes()).hexdigest()


def test_block_clustered_torch_is_bit_identical():
    import torch
    from bmsparse_spgemm_spmv_b200 import generators as G
    for nbr in (40, 300):
        a = G.block_clustered(nbr); b = G.block_clustered_torch(nbr, device="cpu")
        assert a[0] == b[0] and np.array_equal(a[2], b[2].numpy()) and np.array_equal(a[3], b[3].numpy()) and np.array_equal(a[4], b[4].numpy())
