"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def hex_to_u64(lst):
    return np.array([int(x, 16) for x in lst], dtype=np.uint64)


def random_csr(rows, cols, density, seed, empty_block_rows=(), fp16=True):
    """Random CSR with ascending columns per row; some 8-row block rows forced empty."""
    rng = np.random.default_rng(seed)
    mask = rng.random((rows, cols)) < density
    for br in empty_block_rows:
        mask[br * 8:(br + 1) * 8, :] = False
    r, c = np.nonzero(mask)
    v = rng.uniform(-1, 1, r.size).astype(np.float32)
    if fp16:
        v = v.astype(np.float16).astype(np.float32)
        v[v == 0] = 0.25
    rp = np.zeros(rows + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=rows), out=rp[1:])
    return rp.astype(np.int32), c.astype(np.int32), v


def csr_rows(rp):
    return np.repeat(np.arange(rp.size - 1, dtype=np.int32), np.diff(rp))


def assert_structure_equal(got, exp, what=""):
    """got: (keys, bmps, offsets) numpy uint64 from the GPU; exp: OracleMatrix."""
    k, b, o = got
    assert k.size == exp.keys.size, f"{what}: block count {k.size} != {exp.keys.size}"
    assert np.array_equal(k, exp.keys), f"{what}: keys differ"
    assert np.array_equal(b, exp.bmps), f"{what}: bitmaps differ"
    n = min(o.size, exp.offsets.size)
    assert np.array_equal(o[:n], exp.offsets[:n]), f"{what}: offsets differ"


def rel_err(got, exp, floor):
    """max |got-exp| / max(|exp|, floor): the reference's metric (bmSpMatrix.cu:418) with an absolute floor."""
    got = np.asarray(got, np.float64); exp = np.asarray(exp, np.float64)
    if got.size == 0:
        return 0.0
    return float(np.max(np.abs(got - exp) / np.maximum(np.abs(exp), floor)))
