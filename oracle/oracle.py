"""ctypes front-end for the CPU oracle (oracle/bmsp_oracle.c) and the reference's own cusp host
CSR kernels (oracle/_ref/libcusp_ref.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under bmsparse_spgemm_spmv_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_i64 = C.c_int64
_p = C.c_void_p


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_coo_to_bmsp.restype = _i64
        _LIB.orc_bmsp_to_coo.restype = _i64
        _LIB.orc_pair_bitmap.restype = C.c_uint64
        _LIB.orc_pair_bitmap.argtypes = [C.c_uint64, C.c_uint64]
        _LIB.orc_poisson5pt.restype = _i64
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def ref_lib():
    """The reference's cusp host CSR kernels, or None when oracle/_ref was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libcusp_ref.so")
        if not os.path.exists(path):
            return None
        _REF = C.CDLL(path)
        _REF.ref_csr_spgemm.restype = _p
        _REF.ref_result_nnz.restype = C.c_longlong
        _REF.ref_result_nnz.argtypes = [_p]
        _REF.ref_result_copy.argtypes = [_p, _p, _p, _p]
        _REF.ref_result_free.argtypes = [_p]
    return _REF


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_p)


def f16_round(v: np.ndarray) -> np.ndarray:
    """fp32 -> fp16 -> fp32 with the oracle's own rounding (checked against numpy in tests)."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    bits = np.empty(v.shape, dtype=np.uint16)
    lib().orc_f32_to_f16_array(_i64(v.size), _ptr(v), _ptr(bits))
    out = np.empty(v.shape, dtype=np.float32)
    lib().orc_f16_to_f32_array(_i64(v.size), _ptr(bits), _ptr(out))
    return out


@dataclass
class OracleMatrix:
    """keys/bmps/offsets/values exactly as include/bmSpMatrix.h:28-31 holds them (host copies)."""
    num_rows: int
    num_cols: int
    keys: np.ndarray      # uint64 [nblk]
    bmps: np.ndarray      # uint64 [nblk]
    offsets: np.ndarray   # uint64 [nblk] (ingest) or [nblk+1] (product)
    values: np.ndarray    # float32 (fp16-representable when built with f16=True) or float64 (product)
    transposed: bool = False

    @property
    def block_num(self) -> int:
        return int(self.keys.size)

    @property
    def nnz(self) -> int:
        return int(self.values.size)


def coo_to_bmsp(num_rows, num_cols, rows, cols, vals, transposed=False, f16=True) -> OracleMatrix:
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    if f16:
        vals = f16_round(vals)
    n = rows.size
    keys = np.empty(max(n, 1), np.uint64)
    bmps = np.empty(max(n, 1), np.uint64)
    offs = np.empty(max(n, 1), np.uint64)
    vout = np.empty(max(n, 1), np.float32)
    nb = lib().orc_coo_to_bmsp(_i64(n), _ptr(rows), _ptr(cols), _ptr(vals), C.c_int(int(transposed)),
                               _ptr(keys), _ptr(bmps), _ptr(offs), _ptr(vout))
    return OracleMatrix(num_rows, num_cols, keys[:nb].copy(), bmps[:nb].copy(), offs[:nb].copy(),
                        vout[:n].copy(), transposed)


def csr_to_coo(rp, ci):
    rp = np.asarray(rp)
    rows = np.repeat(np.arange(rp.size - 1, dtype=np.int32), np.diff(rp).astype(np.int64))
    return rows, np.asarray(ci, dtype=np.int32)


def csr_to_bmsp(num_rows, num_cols, rp, ci, vals, transposed=False, f16=True) -> OracleMatrix:
    rows, cols = csr_to_coo(rp, ci)
    return coo_to_bmsp(num_rows, num_cols, rows, cols, vals, transposed, f16)


def bmsp_to_coo(m: OracleMatrix):
    n = m.nnz
    rows = np.empty(max(n, 1), np.int32)
    cols = np.empty(max(n, 1), np.int32)
    k = np.ascontiguousarray(m.keys); b = np.ascontiguousarray(m.bmps)
    got = lib().orc_bmsp_to_coo(_i64(m.block_num), _ptr(k), _ptr(b), C.c_int(int(m.transposed)),
                                _ptr(rows), _ptr(cols))
    assert got == n, (got, n)
    return rows[:n].copy(), cols[:n].copy(), np.asarray(m.values).copy()


def spmv(m: OracleMatrix, x: np.ndarray) -> np.ndarray:
    assert not m.transposed
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros(m.num_rows, np.float64)
    k = np.ascontiguousarray(m.keys); b = np.ascontiguousarray(m.bmps); o = np.ascontiguousarray(m.offsets)
    v = np.ascontiguousarray(m.values, dtype=np.float32)
    lib().orc_spmv(C.c_int(m.num_rows), _i64(m.block_num), _ptr(k), _ptr(b), _ptr(o), _ptr(v), _ptr(x), _ptr(y))
    return y


def pair_bitmap(a: int, bt: int) -> int:
    return int(lib().orc_pair_bitmap(C.c_uint64(a), C.c_uint64(bt)))


def spgemm(a: OracleMatrix, bt: OracleMatrix) -> OracleMatrix:
    assert not a.transposed and bt.transposed
    ak, ab, ao = (np.ascontiguousarray(t) for t in (a.keys, a.bmps, a.offsets))
    av = np.ascontiguousarray(a.values, dtype=np.float32)
    bk, bb, bo = (np.ascontiguousarray(t) for t in (bt.keys, bt.bmps, bt.offsets))
    bv = np.ascontiguousarray(bt.values, dtype=np.float32)
    nb = _i64(0); nnz = _i64(0)
    args = [_i64(a.block_num), _ptr(ak), _ptr(ab), _ptr(ao), _ptr(av),
            _i64(bt.block_num), _ptr(bk), _ptr(bb), _ptr(bo), _ptr(bv), C.byref(nb), C.byref(nnz)]
    lib().orc_spgemm(*args, None, None, None, None)
    ck = np.empty(max(nb.value, 1), np.uint64); cb = np.empty(max(nb.value, 1), np.uint64)
    co = np.empty(nb.value + 1, np.uint64); cv = np.empty(max(nnz.value, 1), np.float64)
    lib().orc_spgemm(*args, _ptr(ck), _ptr(cb), _ptr(co), _ptr(cv))
    return OracleMatrix(a.num_rows, bt.num_cols, ck[:nb.value].copy(), cb[:nb.value].copy(), co,
                        cv[:nnz.value].copy(), False)


# ------------------------------------------------------------------ OpenMP variants (full-size parity; pinned to the scalar functions in tests)
def _brp(keys: np.ndarray, nbr: int) -> np.ndarray:
    return np.searchsorted(np.ascontiguousarray(keys) >> np.uint64(32), np.arange(nbr + 1, dtype=np.uint64), side="left").astype(np.int64)


def csr_to_bmsp_omp(num_rows, num_cols, rp, ci, vals, transposed=False, f16=True) -> OracleMatrix:
    rp = np.ascontiguousarray(rp, np.int32); ci = np.ascontiguousarray(ci, np.int32)
    vals = np.ascontiguousarray(vals, np.float32)
    if f16:
        vals = f16_round(vals)
    L = lib(); L.orc_csr_to_bmsp_omp.restype = _i64
    head = [C.c_int(num_rows), _ptr(rp), _ptr(ci), _ptr(vals), C.c_int(int(transposed))]
    nb = L.orc_csr_to_bmsp_omp(*head, None, None, None, None)
    keys = np.empty(max(nb, 1), np.uint64); bmps = np.empty(max(nb, 1), np.uint64); offs = np.empty(max(nb, 1), np.uint64)
    vout = np.empty(max(ci.size, 1), np.float32)
    L.orc_csr_to_bmsp_omp(*head, _ptr(keys), _ptr(bmps), _ptr(offs), _ptr(vout))
    return OracleMatrix(num_rows, num_cols, keys[:nb], bmps[:nb], offs[:nb], vout[:ci.size], transposed)


def spmv_omp(m: OracleMatrix, x: np.ndarray) -> np.ndarray:
    assert not m.transposed
    nbr = (m.num_rows + 7) // 8
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros(m.num_rows, np.float64)
    k = np.ascontiguousarray(m.keys); b = np.ascontiguousarray(m.bmps); o = np.ascontiguousarray(m.offsets)
    v = np.ascontiguousarray(m.values, dtype=np.float32); brp = _brp(k, nbr)
    lib().orc_spmv_omp(C.c_int(m.num_rows), C.c_int(nbr), _ptr(brp), _ptr(k), _ptr(b), _ptr(o), _ptr(v), _ptr(x), _ptr(y))
    return y


def spgemm_omp(a: OracleMatrix, bt: OracleMatrix) -> OracleMatrix:
    """orc_spgemm, one A block row per OpenMP task; values fp32 (rounded from the double sums)."""
    assert not a.transposed and bt.transposed
    nbr = (a.num_rows + 7) // 8
    ak, ab, ao = (np.ascontiguousarray(t) for t in (a.keys, a.bmps, a.offsets))
    av = np.ascontiguousarray(a.values, dtype=np.float32)
    bk, bb, bo = (np.ascontiguousarray(t) for t in (bt.keys, bt.bmps, bt.offsets))
    bv = np.ascontiguousarray(bt.values, dtype=np.float32)
    brp = _brp(ak, nbr)
    rb = np.zeros(nbr + 1, np.int64); rn = np.zeros(nbr + 1, np.int64)
    head = [C.c_int(nbr), _ptr(brp), _ptr(ak), _ptr(ab), _ptr(ao), _ptr(av), _i64(bt.block_num), _ptr(bk), _ptr(bb), _ptr(bo), _ptr(bv),
            _ptr(rb), _ptr(rn)]
    if lib().orc_spgemm_omp(*head, None, None, None, None):
        raise MemoryError("orc_spgemm_omp")
    nb, nnz = int(rb[nbr]), int(rn[nbr])
    ck = np.empty(max(nb, 1), np.uint64); cb = np.empty(max(nb, 1), np.uint64); co = np.empty(nb + 1, np.uint64); cv = np.empty(max(nnz, 1), np.float32)
    if lib().orc_spgemm_omp(*head, _ptr(ck), _ptr(cb), _ptr(co), _ptr(cv)):
        raise MemoryError("orc_spgemm_omp")
    return OracleMatrix(a.num_rows, bt.num_cols, ck[:nb], cb[:nb], co, cv[:nnz], False)


# ------------------------------------------------------------------ cusp host CSR kernels
def csr_spmv(rp, ci, v, x, threads=1) -> np.ndarray:
    rp = np.ascontiguousarray(rp, np.int32); ci = np.ascontiguousarray(ci, np.int32)
    v = np.ascontiguousarray(v, np.float32); x = np.ascontiguousarray(x, np.float32)
    y = np.empty(rp.size - 1, np.float32)
    lib().orc_csr_spmv(C.c_int(rp.size - 1), _ptr(rp), _ptr(ci), _ptr(v), _ptr(x), _ptr(y), C.c_int(threads))
    return y


def csr_spgemm(a_rows, b_cols, a_rp, a_ci, a_v, b_rp, b_ci, b_v, drop_zeros=True, threads=1):
    a_rp = np.ascontiguousarray(a_rp, np.int32); a_ci = np.ascontiguousarray(a_ci, np.int32)
    a_v = np.ascontiguousarray(a_v, np.float32)
    b_rp = np.ascontiguousarray(b_rp, np.int32); b_ci = np.ascontiguousarray(b_ci, np.int32)
    b_v = np.ascontiguousarray(b_v, np.float32)
    c_rp = np.empty(a_rows + 1, np.int32); nnz = _i64(0)
    base = [C.c_int(a_rows), C.c_int(b_cols), _ptr(a_rp), _ptr(a_ci), _ptr(a_v), _ptr(b_rp), _ptr(b_ci), _ptr(b_v),
            C.byref(nnz), _ptr(c_rp)]
    lib().orc_csr_spgemm(*base, None, None, C.c_int(int(drop_zeros)), C.c_int(threads))
    c_ci = np.empty(max(nnz.value, 1), np.int32); c_v = np.empty(max(nnz.value, 1), np.float32)
    lib().orc_csr_spgemm(*base, _ptr(c_ci), _ptr(c_v), C.c_int(int(drop_zeros)), C.c_int(threads))
    return c_rp, c_ci[:nnz.value].copy(), c_v[:nnz.value].copy()


def ref_csr_spmv(rows, cols, rp, ci, v, x, omp=False) -> np.ndarray:
    r = ref_lib()
    rp = np.ascontiguousarray(rp, np.int32); ci = np.ascontiguousarray(ci, np.int32)
    v = np.ascontiguousarray(v, np.float32); x = np.ascontiguousarray(x, np.float32)
    y = np.empty(rows, np.float32)
    (r.ref_csr_spmv_omp if omp else r.ref_csr_spmv_seq)(C.c_int(rows), C.c_int(cols), _ptr(rp), _ptr(ci), _ptr(v), _ptr(x), _ptr(y))
    return y


def ref_csr_spgemm(a_rows, a_cols, a_rp, a_ci, a_v, b_rows, b_cols, b_rp, b_ci, b_v, omp=False, copy=True):
    r = ref_lib()
    arrs = [np.ascontiguousarray(t, d) for t, d in ((a_rp, np.int32), (a_ci, np.int32), (a_v, np.float32),
                                                     (b_rp, np.int32), (b_ci, np.int32), (b_v, np.float32))]
    h = r.ref_csr_spgemm(C.c_int(a_rows), C.c_int(a_cols), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                         C.c_int(b_rows), C.c_int(b_cols), _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5]),
                         C.c_int(int(omp)))
    h = _p(h)
    n = r.ref_result_nnz(h)
    out = None
    if copy:
        rp = np.empty(a_rows + 1, np.int32); ci = np.empty(max(n, 1), np.int32); v = np.empty(max(n, 1), np.float32)
        r.ref_result_copy(h, _ptr(rp), _ptr(ci), _ptr(v))
        out = (rp, ci[:n].copy(), v[:n].copy())
    r.ref_result_free(h)
    return out if copy else n


def poisson5pt(m: int, n: int):
    N = m * n
    rp = np.empty(N + 1, np.int32); ci = np.empty(5 * N, np.int32); v = np.empty(5 * N, np.float32)
    nnz = lib().orc_poisson5pt(C.c_int(m), C.c_int(n), _ptr(rp), _ptr(ci), _ptr(v))
    return rp, ci[:nnz].copy(), v[:nnz].copy()


def max_threads() -> int:
    return int(lib().orc_max_threads())


def use_all_cores() -> int:
    """Undo an inherited OMP_NUM_THREADS=1 (torch.distributed.run exports it to every worker): the CPU baselines run on every
    core the process may use.  Returns the OpenMP thread count now in effect."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_threads.restype = C.c_int
    return int(lib().orc_set_threads(C.c_int(n)))


def read_mtx(path):
    """Minimal MatrixMarket coordinate reader with the semantics of src/bmSpMatrix.cu:111-159:
    'symmetric' in the banner mirrors off-diagonals; indices 1-based; values parsed as double."""
    with open(path) as f:
        first = f.readline()
        sym = "symmetric" in first
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        nr, nc, nl = (int(t) for t in line.split())
        r = []; c = []; v = []
        for _ in range(nl):
            t = f.readline().split()
            i, j = int(t[0]) - 1, int(t[1]) - 1
            x = float(t[2]) if len(t) > 2 else 1.0
            r.append(i); c.append(j); v.append(x)
            if sym and i != j:
                r.append(j); c.append(i); v.append(x)
    return nr, nc, np.array(r, np.int32), np.array(c, np.int32), np.array(v, np.float64)


# ------------------------------------------------------------------ the reference's own CUDA operators (GPU box only)
def ref_cuda_bin(name: str):
    """oracle/_ref/ref_spgemm | ref_spmv: the reference's bmSparse_mult / bmSparse_SpMV compiled for sm_100 from
    /root/reference by `make -C oracle refcuda`; None when they were never built."""
    p = os.path.join(_HERE, "_ref", name)
    return p if os.path.exists(p) else None


def write_bin(path, m: OracleMatrix, value_dtype):
    v = np.ascontiguousarray(m.values, dtype=value_dtype)
    with open(path, "wb") as f:
        np.array([m.num_rows, m.num_cols, m.block_num, v.size, v.itemsize], np.int64).tofile(f)
        np.ascontiguousarray(m.keys, np.uint64).tofile(f)
        np.ascontiguousarray(m.bmps, np.uint64).tofile(f)
        np.ascontiguousarray(m.offsets[:m.block_num], np.uint64).tofile(f)
        v.tofile(f)


def run_ref_spgemm(a: OracleMatrix, bt: OracleMatrix, workdir, tc_version=5, mode=0, reps=1, keep=True):
    """Run the reference's bmSparse_mult<half,float> on the GPU.  Returns (C as OracleMatrix fp32, best microseconds)."""
    exe = ref_cuda_bin("ref_spgemm")
    pa, pb, pc = (os.path.join(workdir, n) for n in ("refA.bin", "refB.bin", "refC.bin"))
    write_bin(pa, a, np.float16); write_bin(pb, bt, np.float16)
    out = subprocess.run([exe, pa, pb, pc, str(tc_version), str(mode), str(reps)], capture_output=True, text=True, timeout=1800)
    line = [l for l in out.stdout.splitlines() if l.startswith("REF_SPGEMM_US")]
    if out.returncode != 0 or not line:
        raise RuntimeError(f"reference spgemm failed: rc={out.returncode}\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    tok = line[0].split()
    us, status = float(tok[1]), int(tok[7])
    if status != 0:
        raise RuntimeError(f"reference spgemm left CUDA error {status}")
    with open(pc, "rb") as f:
        h = np.fromfile(f, np.int64, 5)
        k = np.fromfile(f, np.uint64, h[2]); b = np.fromfile(f, np.uint64, h[2]); o = np.fromfile(f, np.uint64, h[4])
        v = np.fromfile(f, np.float32, h[3])
    if not keep:
        for p in (pa, pb, pc):
            os.remove(p)
    return OracleMatrix(int(h[0]), int(h[1]), k, b, o, v, False), us


def sha256_u64(arr) -> str:
    """SHA-256 of the little-endian bytes of a uint64 array (what oracle/ref_bmsparse_driver.cu prints for the reference's arrays)."""
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(arr, dtype="<u8").tobytes() if arr.size < (1 << 24) else memoryview(np.ascontiguousarray(arr, dtype="<u8"))).hexdigest()


def run_ref_spgemm_digest(a: OracleMatrix, bt: OracleMatrix, workdir, tc_version=5, mode=0, reps=1, sample_blocks=0, reuse_inputs=False):
    """Run the reference's bmSparse_mult<half,float> without dumping C (sample_blocks = 0) or dumping three windows of
    `sample_blocks` C blocks (first / middle / last).  Returns a dict: us, c_blocks, c_nnz, sha256 {keys, bmps, offsets},
    offsets_len, and `windows`: [{first_block, keys, bmps, offsets, first_value, values}] when sampled."""
    exe = ref_cuda_bin("ref_spgemm")
    pa, pb, pc = (os.path.join(workdir, n) for n in ("refA.bin", "refB.bin", "refCs.bin"))
    if not (reuse_inputs and os.path.exists(pa) and os.path.exists(pb)):
        write_bin(pa, a, np.float16); write_bin(pb, bt, np.float16)
    out_path = pc if sample_blocks > 0 else "-"
    out = subprocess.run([exe, pa, pb, out_path, str(tc_version), str(mode), str(reps), str(int(sample_blocks))], capture_output=True, text=True, timeout=1800)
    line = [l for l in out.stdout.splitlines() if l.startswith("REF_SPGEMM_US")]
    sha = [l for l in out.stdout.splitlines() if l.startswith("REF_SPGEMM_SHA256")]
    if out.returncode != 0 or not line or not sha:
        raise RuntimeError(f"reference spgemm failed: rc={out.returncode}\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    tok = line[0].split()
    if int(tok[7]) != 0:
        raise RuntimeError(f"reference spgemm left CUDA error {tok[7]}")
    st = sha[0].split()
    res = {"us": float(tok[1]), "c_blocks": int(tok[3]), "c_nnz": int(tok[5]),
           "sha256": {"keys": st[2], "bmps": st[4], "offsets": st[6]}, "offsets_len": int(st[8]), "windows": []}
    if sample_blocks > 0:
        with open(pc, "rb") as f:
            h = np.fromfile(f, np.int64, 5)
            if h[2] >= 0:       # small product: the driver dumped everything
                nb = int(h[2])
                k = np.fromfile(f, np.uint64, nb); b = np.fromfile(f, np.uint64, nb); o = np.fromfile(f, np.uint64, int(h[4])); v = np.fromfile(f, np.float32, int(h[3]))
                res["windows"].append({"first_block": 0, "keys": k, "bmps": b, "offsets": o, "first_value": 0, "values": v})
            else:
                for _ in range(int(h[4])):
                    wh = np.fromfile(f, np.int64, 4)
                    n = int(wh[1])
                    k = np.fromfile(f, np.uint64, n); b = np.fromfile(f, np.uint64, n); o = np.fromfile(f, np.uint64, n + 1)
                    v = np.fromfile(f, np.float32, int(wh[3]))
                    res["windows"].append({"first_block": int(wh[0]), "keys": k, "bmps": b, "offsets": o, "first_value": int(wh[2]), "values": v})
        os.remove(pc)
    return res


def run_ref_spmv(a: OracleMatrix, workdir, reps=1):
    """Run the reference's bmSparse_SpMV<float,float> (x = ones, as its main does).  Returns (y, best microseconds)."""
    exe = ref_cuda_bin("ref_spmv")
    pa, py = os.path.join(workdir, "refA32.bin"), os.path.join(workdir, "refY.bin")
    write_bin(pa, a, np.float32)
    out = subprocess.run([exe, pa, py, str(reps)], capture_output=True, text=True, timeout=1800)
    line = [l for l in out.stdout.splitlines() if l.startswith("REF_SPMV_US")]
    if out.returncode != 0 or not line:
        raise RuntimeError(f"reference spmv failed: rc={out.returncode}\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    us = float(line[0].split()[1])
    with open(py, "rb") as f:
        n = int(np.fromfile(f, np.int64, 1)[0]); y = np.fromfile(f, np.float32, n)
    return y, us
