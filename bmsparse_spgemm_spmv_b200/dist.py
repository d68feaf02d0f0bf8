"""Block-row sharding over one process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

The reference is single-GPU (no collectives anywhere); this is the multi-GPU layer the
north star adds.  Both operators shard by block rows and need no reduction:

  SpMV    rank r owns a contiguous range of block rows and the matching slice of x / y.  Before each
          product it fetches only the slices of x its block columns touch ("range halo": for a banded
          matrix two neighbours, for a scattered one everybody -- then it degenerates to an all-gather).
          y is written straight into the owner's slice of the next x, so repeated products ping-pong
          between two buffers with no copy.
  SpGEMM  rank r multiplies its block rows of A by a replicated B^t (bmSparse_mult(..., brow_range)); the
          concatenation of the per-rank C arrays is bit-identical to the single-GPU product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def split_by_weight(weights: np.ndarray, nparts: int, align: int = 1) -> np.ndarray:
    """Contiguous split of len(weights) items into nparts ranges of near-equal total weight.
    Returns int64 bounds[nparts+1]; interior bounds are multiples of `align`."""
    n = len(weights)
    prefix = np.concatenate([[0], np.cumsum(np.asarray(weights, np.float64))])
    bounds = np.zeros(nparts + 1, np.int64)
    bounds[nparts] = n
    for p in range(1, nparts):
        b = int(np.searchsorted(prefix, prefix[-1] * p / nparts, side="left"))
        b = (b // align) * align
        bounds[p] = min(max(b, bounds[p - 1]), n)
    return bounds


def csr_row_slice(rp, ci, v, r0, r1):
    """rows [r0, r1) of a CSR triple, row_ptr rebased to 0 (columns untouched)."""
    s, e = int(rp[r0]), int(rp[r1])
    return (np.asarray(rp[r0:r1 + 1]) - s).astype(np.int32), ci[s:e], v[s:e]


class ShardedSpMV:
    """y = A x with A sharded by rows and x sharded the same way (square matrices).

    Parameters
    ----------
    row_bounds : global row split (len world+1), interior bounds multiples of 8
    local_csr  : (row_ptr, col_idx, vals) of this rank's rows, GLOBAL column indices, ascending per row
    spmv_fn    : callable(x_ext, y_out) computing the local product; default builds a bmSpMatrix on the
                 current CUDA device and calls bmSparse_SpMV (the CPU tests inject the oracle here)
    """

    def __init__(self, row_bounds, local_csr, n_cols, group=None, device=None, spmv_fn=None, build_fn=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.bounds = np.asarray(row_bounds, np.int64)
        self.device = device if device is not None else torch.device("cpu")
        rp, ci, v = local_csr
        self.own_lo, self.own_hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        lo = min(self.own_lo, int(ci.min()) if len(ci) else self.own_lo)
        hi = max(self.own_hi, int(ci.max()) + 1 if len(ci) else self.own_hi)
        self.ext_lo, self.ext_hi = (lo // 8) * 8, min(((hi + 7) // 8) * 8, n_cols)
        # everybody learns everybody's extended range
        mine = torch.tensor([self.ext_lo, self.ext_hi], dtype=torch.int64, device=self.device)
        allr = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allr, mine, group=group)
        ranges = [tuple(int(t) for t in a.cpu()) for a in allr]
        # what I receive from peer p: overlap(my ext range, p's owned range); what I send: the mirror image
        self.recv, self.send = [], []
        for p in range(self.world):
            if p == self.rank:
                continue
            plo, phi = int(self.bounds[p]), int(self.bounds[p + 1])
            a, b = max(self.ext_lo, plo), min(self.ext_hi, phi)
            if a < b:
                self.recv.append((p, a, b))
            a, b = max(ranges[p][0], self.own_lo), min(ranges[p][1], self.own_hi)
            if a < b:
                self.send.append((p, a, b))
        n_ext = self.ext_hi - self.ext_lo
        self.x = [torch.zeros(n_ext, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.cur = 0
        self.halo_bytes = sum(b - a for _, a, b in self.recv) * 4
        local_ci = (np.asarray(ci, np.int64) - self.ext_lo).astype(np.int32)
        n_rows = self.own_hi - self.own_lo
        if spmv_fn is not None:
            self._spmv = spmv_fn
            self.local = build_fn(n_rows, n_ext, rp, local_ci, v) if build_fn else None
        else:
            from .matrix import bmSpMatrix
            from .ops import bmSparse_SpMV
            self.local = bmSpMatrix.from_csr(n_rows, n_ext, rp, local_ci, v)
            self._spmv = lambda x_ext, y_out: bmSparse_SpMV(self.local, x_ext, y_out)

    def own_slice(self, buf):
        return buf[self.own_lo - self.ext_lo: self.own_hi - self.ext_lo]

    def set_x(self, x_own: torch.Tensor):
        self.own_slice(self.x[self.cur]).copy_(x_own)

    def exchange(self):
        """fetch the halo slices of the current x from their owners (NCCL send/recv over NVLink)."""
        buf = self.x[self.cur]
        ops = []
        for p, a, b in self.send:
            ops.append(dist.P2POp(dist.isend, buf[a - self.ext_lo: b - self.ext_lo], p, group=self.group))
        for p, a, b in self.recv:
            ops.append(dist.P2POp(dist.irecv, buf[a - self.ext_lo: b - self.ext_lo], p, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def step(self):
        """one product: halo exchange of x, then y = A_local x_ext written into the next x's own slice."""
        self.exchange()
        nxt = 1 - self.cur
        self._spmv(self.x[self.cur], self.own_slice(self.x[nxt]))
        self.cur = nxt

    def y_own(self):
        return self.own_slice(self.x[self.cur])


def broadcast_matrix(M, src: int = 0, group=None):
    """Replicate a bmSpMatrix (e.g. B^t for SpGEMM) from rank `src` to every rank: four array broadcasts."""
    from .matrix import bmSpMatrix
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if rank == src:
        v = M._view()
        hdr = torch.tensor([v.num_rows, v.num_cols, v.block_num, v.nnz, v.offsets_len, v.dtype, v.transposed], dtype=torch.int64, device=dev)
    else:
        hdr = torch.zeros(7, dtype=torch.int64, device=dev)
    dist.broadcast(hdr, src, group=group)
    nr, nc, nblk, nnz, noff, dt, tr = (int(t) for t in hdr.cpu())
    vt = torch.float16 if dt == 0 else torch.float32
    if rank == src:
        arrs = [M.keys.clone(), M.bmps.clone(), M.offsets.clone(), M.values.clone()]
    else:
        arrs = [torch.empty(nblk, dtype=torch.int64, device=dev), torch.empty(nblk, dtype=torch.int64, device=dev),
                torch.empty(noff, dtype=torch.int64, device=dev), torch.empty(nnz, dtype=vt, device=dev)]
    for a in arrs:
        dist.broadcast(a, src, group=group)
    if rank == src:
        return M
    return bmSpMatrix.from_arrays(nr, nc, nblk, arrs[0], arrs[1], arrs[2], arrs[3], transpose=bool(tr))
