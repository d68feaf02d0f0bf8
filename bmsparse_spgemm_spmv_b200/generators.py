"""Seeded synthetic inputs for the BASELINE.json configs (host side, numpy; not on the hot path).

All randomness is counter-based splitmix64(seed, index), so a given (generator, seed, shape) is the
same everywhere.  Every generator returns CSR (row_ptr int32, col_idx int32 ascending per row,
vals float32 already rounded to fp16-representable values) plus the shape.

  poisson5pt(m, n)          cusp::gallery::poisson5pt (cusp/gallery/detail/poisson.inl:28-46): index x + m*y,
                            diagonal 4, neighbours -1                                   -> configs 1, 2
  uniform_random(n, k)      k distinct uniform columns per row                           -> config 3
  block_clustered(nbr, ...) 8x8 blocks in a +-16 block band, 30 % present, 50 % fill      -> config 4
  rmat(scale, ef)           R-MAT (0.57, 0.19, 0.19, 0.05), duplicates merged             -> config 5
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """splitmix64 of (seed + (idx+1) * golden) -- vectorised, uint64 in / uint64 out."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _unit(seed: int, idx: np.ndarray) -> np.ndarray:
    """uniform [0,1) float64 from the top 53 bits."""
    return (splitmix64(seed, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def values_fp16(seed: int, n: int) -> np.ndarray:
    """uniform(-1,1) rounded to fp16, returned as float32 (zero avoided so structure == value pattern)."""
    v = (_unit(seed ^ 0x5EED, np.arange(n, dtype=np.uint64)) * 2.0 - 1.0).astype(np.float16)
    v[v == 0] = np.float16(0.5)
    return v.astype(np.float32)


def poisson5pt(m: int, n: int):
    N = m * n
    i = np.arange(N, dtype=np.int64)
    x = i % m
    y = i // m
    cols = np.stack([i - m, i - 1, i, i + 1, i + m], axis=1)
    valid = np.stack([y > 0, x > 0, np.ones(N, bool), x < m - 1, y < n - 1], axis=1)
    vals = np.broadcast_to(np.array([-1, -1, 4, -1, -1], np.float32), (N, 5))
    rp = np.zeros(N + 1, np.int64)
    np.cumsum(valid.sum(1), out=rp[1:])
    return N, N, rp.astype(np.int32), cols[valid].astype(np.int32), vals[valid].copy()


def uniform_random(n: int, k: int, seed: int = 2, ncols: int | None = None):
    ncols = ncols or n
    idx = np.arange(n * k, dtype=np.uint64)
    c = (splitmix64(seed, idx) % np.uint64(ncols)).astype(np.int64).reshape(n, k)
    salt = 1
    while True:
        c.sort(axis=1)
        dup = np.zeros_like(c, bool)
        dup[:, 1:] = c[:, 1:] == c[:, :-1]
        nd = int(dup.sum())
        if nd == 0:
            break
        where = np.flatnonzero(dup.ravel()).astype(np.uint64)
        c.ravel()[where] = (splitmix64(seed + 7919 * salt, where) % np.uint64(ncols)).astype(np.int64)
        salt += 1
    rp = (np.arange(n + 1, dtype=np.int64) * k).astype(np.int32)
    return n, ncols, rp, c.ravel().astype(np.int32), values_fp16(seed, n * k)


def block_clustered(nbr: int, half_band: int = 16, p_block: float = 0.30, seed: int = 3):
    """n = 8*nbr rows.  Block (I,J) is a candidate iff J in [I-half_band, I+half_band-1]; present with
    probability p_block (diagonal always); each of its 64 cells present with probability 0.5 (>=1 forced)."""
    W = 2 * half_band
    I = np.arange(nbr, dtype=np.int64)[:, None]
    d = np.arange(W, dtype=np.int64)[None, :]
    J = I + d - half_band
    lin = (I * W + d).astype(np.uint64)
    present = (_unit(seed, lin) < p_block) | (J == I)
    present &= (J >= 0) & (J < nbr)
    mask = splitmix64(seed + 101, lin)
    mask = np.where(mask == 0, np.uint64(1), mask)
    mask = np.where(present, mask, np.uint64(0))
    # byte ri (MSB first) of each mask = row ri of the block; bit (MSB first) = column
    be = mask.astype(">u8").view(np.uint8).reshape(nbr, W, 8)          # [I, d, ri]
    rows_bytes = np.ascontiguousarray(be.transpose(0, 2, 1))           # [I, ri, d]
    bits = np.unpackbits(rows_bytes, axis=2)                           # [I, ri, d*8 + ci]  (MSB first)
    nnz_per_row = bits.sum(axis=2).reshape(-1)
    rp = np.zeros(nbr * 8 + 1, np.int64)
    np.cumsum(nnz_per_row, out=rp[1:])
    flat = np.flatnonzero(bits.reshape(-1))
    rowi = flat // (W * 8)
    col = (rowi // 8 - half_band) * 8 + flat % (W * 8)
    n = nbr * 8
    return n, n, rp.astype(np.int32), col.astype(np.int32), values_fp16(seed, flat.size)


def rmat(scale: int, edge_factor: int = 16, a=0.57, b=0.19, c=0.19, seed: int = 4):
    n = 1 << scale
    ne = edge_factor * n
    e = np.arange(ne, dtype=np.uint64)
    r = np.zeros(ne, np.int64)
    col = np.zeros(ne, np.int64)
    for lvl in range(scale):
        u = _unit(seed + 1000003 * (lvl + 1), e)
        rbit = u >= (a + b)
        cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        r = (r << 1) | rbit
        col = (col << 1) | cbit
    key = np.unique((r << 32) | col)
    r = key >> 32
    col = key & 0xFFFFFFFF
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=rp[1:])
    return n, n, rp.astype(np.int32), col.astype(np.int32), values_fp16(seed, key.size)


# ---- the same counter-based generators evaluated with torch (on the GPU for the big configs: rmat(22) takes minutes in numpy) ----
def _t_lsr(z, k):
    """logical shift right of int64 bit patterns"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _t_i64(c: int) -> int:
    """uint64 constant as the int64 with the same bits"""
    return c - (1 << 64) if c >= (1 << 63) else c


def splitmix64_torch(seed: int, idx):
    """splitmix64() above on torch int64 tensors (two's-complement wrap-around == uint64 arithmetic); returns int64 bit patterns."""
    z = (idx + 1) * _t_i64(0x9E3779B97F4A7C15) + _t_i64(seed & 0xFFFFFFFFFFFFFFFF)
    z = (z ^ _t_lsr(z, 30)) * _t_i64(0xBF58476D1CE4E5B9)
    z = (z ^ _t_lsr(z, 27)) * _t_i64(0x94D049BB133111EB)
    return z ^ _t_lsr(z, 31)


def _unit_torch(seed: int, idx):
    import torch
    return _t_lsr(splitmix64_torch(seed, idx), 11).to(torch.float64) * (1.0 / (1 << 53))


def values_fp16_torch(seed: int, n: int, device):
    import torch
    v = _unit_torch(seed ^ 0x5EED, torch.arange(n, dtype=torch.int64, device=device)) * 2.0 - 1.0
    # float64 -> fp16 correctly rounded (numpy's direct conversion; torch goes through float32 and double-rounds): 11 significant
    # bits for normal numbers, a 2^-24 grid below 2^-14, ties to even -- all exact in float64
    m, e = torch.frexp(v)
    q = torch.where(v.abs() >= 2.0 ** -14, torch.ldexp(torch.round(torch.ldexp(m, torch.tensor(11, device=device))), e - 11),
                    torch.round(v * 2.0 ** 24) * 2.0 ** -24)
    q = q.to(torch.float32)
    q[q == 0] = 0.5
    return q


def rmat_torch(scale: int, edge_factor: int = 16, a=0.57, b=0.19, c=0.19, seed: int = 4, device="cuda"):
    """rmat() above, bit-identical, as torch tensors on `device`: (n, n, row_ptr int32, col_idx int32, vals float32)."""
    import torch
    n = 1 << scale
    ne = edge_factor * n
    e = torch.arange(ne, dtype=torch.int64, device=device)
    r = torch.zeros(ne, dtype=torch.int64, device=device)
    col = torch.zeros(ne, dtype=torch.int64, device=device)
    for lvl in range(scale):
        u = _unit_torch(seed + 1000003 * (lvl + 1), e)
        rbit = u >= (a + b)
        cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        r = (r << 1) | rbit
        col = (col << 1) | cbit
        del u, rbit, cbit
    key = torch.unique((r << 32) | col)
    del r, col, e
    rows = key >> 32
    rp = torch.zeros(n + 1, dtype=torch.int64, device=device)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return n, n, rp.to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32), values_fp16_torch(seed, key.numel(), device)


def block_clustered_torch(nbr: int, half_band: int = 16, p_block: float = 0.30, seed: int = 3, device="cuda"):
    """block_clustered() above, bit-identical, as torch tensors on `device` (the 4M-row config takes 30 s in numpy)."""
    import torch
    W = 2 * half_band
    I = torch.arange(nbr, dtype=torch.int64, device=device)[:, None]
    d = torch.arange(W, dtype=torch.int64, device=device)[None, :]
    J = I + d - half_band
    lin = I * W + d
    present = (_unit_torch(seed, lin) < p_block) | (J == I)
    present &= (J >= 0) & (J < nbr)
    mask = splitmix64_torch(seed + 101, lin)
    mask = torch.where(mask == 0, torch.ones_like(mask), mask)
    mask = torch.where(present, mask, torch.zeros_like(mask))
    del present, lin, J
    # byte ri (MSB first) of each mask = row ri of the block; bit (MSB first) = column
    shifts = (56 - 8 * torch.arange(8, dtype=torch.int64, device=device))[None, :, None]         # [1, ri, 1]
    rows_bytes = (_t_lsr(mask[:, None, :], 0) >> shifts) & 0xFF                                      # [I, ri, d]
    del mask
    bitpos = (7 - torch.arange(8, dtype=torch.int64, device=device))[None, None, None, :]
    bits = ((rows_bytes[..., None] >> bitpos) & 1).to(torch.bool).reshape(nbr * 8, W * 8)           # [I*8 + ri, d*8 + ci]
    del rows_bytes
    n = nbr * 8
    rp = torch.zeros(n + 1, dtype=torch.int64, device=device)
    rp[1:] = torch.cumsum(bits.sum(dim=1), 0)
    nz = torch.nonzero(bits)                                                                         # row-major, like np.flatnonzero
    del bits
    rowi, within = nz[:, 0], nz[:, 1]
    col = (rowi // 8 - half_band) * 8 + within
    return n, n, rp.to(torch.int32), col.to(torch.int32), values_fp16_torch(seed, int(col.numel()), device)


def x_vector(n: int, seed: int = 1) -> np.ndarray:
    return (_unit(seed ^ 0xABCD, np.arange(n, dtype=np.uint64)) * 2.0 - 1.0).astype(np.float32)
