"""Strong-scaling measurements on one box (the same matrix at every N; one process per GPU): used by bench.py's "strong" section and
by tools/rmat_scale.py.

  rmat_spmv    BASELINE config 5: R-MAT (a,b,c,d = .57,.19,.19,.05; 16 edges/vertex, duplicates merged) SpMV.  Rows split by SpMV
               cost (bmsp_partition_block_rows); every rank keeps x for its extended column range, the peers' slices arrive over
               NVLink peer memory (bmsp_spmv_halo; for R-MAT every rank needs every slice, so the exchange is an all-gather written
               by the producers).  value = algorithmic bytes of the WHOLE matrix / max-over-ranks device time per product.
  rmat_spgemm  A*A: A's block rows split by candidate pairs into N x k chunks dealt to the ranks block-cyclically; B^t replicated;
               chunks of <= chunk_pairs candidate pairs (the scale-22 product, ~7e10 values, fits no GPU: each chunk's C is reduced
               to a checksum -- blocks, values, sum of keys, sum of values -- and dropped).  The checksums are summed over ranks
               and are independent of N.  value = 2 * scalar products / max-over-ranks time of the bmsp_spgemm calls.
  poisson_split  the P4096 grid (BASELINE config 2) cut into N slabs of 4096/N grid rows, halo exchange fused into the kernel.
"""
from __future__ import annotations

import time

import numpy as np


class Ctx:
    def __init__(self, B, G, torch, dist, rank, world, dev):
        self.B, self.G, self.torch, self.dist, self.rank, self.world, self.dev = B, G, torch, dist, rank, world, dev

    def allmax(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, vs, dtype=None):
        t = self.torch.tensor(vs, dtype=dtype or self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().tolist()

    def agree(self, ok: bool) -> bool:
        """every rank takes rank 0's decision (time budgets differ by a few ms between ranks; a collective section must not split)"""
        if self.world == 1:
            return ok
        t = self.torch.tensor([1 if ok else 0], device=self.dev)
        self.dist.broadcast(t, 0)
        return bool(int(t.item()))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()


def time_steps(c: Ctx, step, steps, warmup=5):
    torch = c.torch
    for _ in range(warmup):
        step()
    c.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); e1.synchronize()
    return c.allmax(e0.elapsed_time(e1) / steps)


def sharded_spmv(c: Ctx, A, n, rp, ci, v, x0, steps, halo="auto", diagnostics=False):
    """times y = A x repeated (x <- y) with A row-sharded over the ranks.  A: the whole matrix on this GPU (for the partition and,
    at N = 1, the product itself); rp/ci/v: host CSR of the whole matrix; returns (ms per product, info)."""
    from bmsparse_spgemm_spmv_b200.dist import ShardedSpMV, csr_row_slice
    B, torch = c.B, c.torch
    d = lambda a: torch.from_numpy(a).to(c.dev)
    if c.world == 1:
        x = d(x0); y = torch.empty(n, device=c.dev)
        ms = time_steps(c, lambda: B.bmSparse_SpMV(A, x, y), steps)
        return ms, {"halo": "none"}
    if A.nnz < 2.5 * A.block_num:
        # scattered matrix: split by measured cost (local product + the all-gather pushes), three rounds at set-up time
        from bmsparse_spgemm_spmv_b200.dist import balanced_block_row_bounds
        bounds = balanced_block_row_bounds(A).astype(np.int64) * 8
    else:
        bounds = A.partition_block_rows(c.world).astype(np.int64) * 8
    bounds[-1] = n
    lcsr = csr_row_slice(rp, ci, v, int(bounds[c.rank]), int(bounds[c.rank + 1]))
    sh = ShardedSpMV(bounds, lcsr, n, device=c.dev, halo=halo)
    sh.set_x(d(x0[bounds[c.rank]:bounds[c.rank + 1]]))
    ms = time_steps(c, sh.step, steps)
    sh.check()
    info = {"halo": "peer-memory" if sh.p2p is not None else "nccl", "halo_bytes_in_per_rank": sh.halo_bytes}
    if diagnostics:
        xl = sh.x[sh.cur]; yl = torch.empty(sh.own_hi - sh.own_lo, device=c.dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            B.bmSparse_SpMV(sh.local, xl, yl)
        e0.record()
        for _ in range(steps):
            B.bmSparse_SpMV(sh.local, xl, yl)
        e1.record(); e1.synchronize()
        mine = {"rank": c.rank, "local_ms": round(e0.elapsed_time(e1) / steps, 4), "rows": sh.own_hi - sh.own_lo, "blocks": sh.local.block_num}
        per = [None] * c.world
        c.dist.all_gather_object(per, mine)
        info["per_rank"] = per
    sh.close()
    return ms, info


def rmat_spgemm(c: Ctx, A, Bt, rp, ci, chunk_pairs=3e9, max_chunks=0):
    B, torch = c.B, c.torch
    rowlen = np.diff(rp).astype(np.int64)
    flops = 2 * int(rowlen[ci].sum())
    bcol = A.block_col.cpu().numpy()
    blen = np.diff(Bt.block_row_ptr.cpu().numpy().astype(np.int64))
    cand_total = int(blen[bcol].sum())
    cpr = max(1, int(np.ceil(cand_total / c.world / chunk_pairs)))
    bounds = A.partition_block_rows(c.world * cpr, Bt)
    # chunks dealt in rounds of N, every other round in reverse rank order: the hub rows sit in the first chunks and a hub chunk costs
    # more per candidate pair than a tail chunk, so contiguous ranges would leave rank 0 with the slowest ones, and a plain cyclic deal
    # still hands rank 0 the heavier chunk of every round
    from bmsparse_spgemm_spmv_b200.dist import deal_chunks
    mine = deal_chunks(c.world * cpr, c.world, c.rank)
    c.barrier()
    spent = 0.0; blocks = 0; nnz = 0; keysum = 0; valsum = 0.0; cand = 0; surv = 0; done = 0; err = None
    for ch in mine:
        r0, r1 = int(bounds[ch]), int(bounds[ch + 1])
        if r1 <= r0:
            continue
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            C, info = B.bmSparse_mult(A, Bt, None, 0, False, 5, brow_range=(r0, r1))
        except B.BmspError as e:
            err = f"rows [{r0},{r1}): {e}"
            break
        e1.record(); e1.synchronize(); spent += e0.elapsed_time(e1) * 1e-3          # device time of the call (it syncs internally)
        blocks += C.block_num; nnz += C.nnz; cand += info.candidate_pairs; surv += info.surviving_pairs
        if C.block_num:
            keysum = (keysum + int(C.keys.sum().item())) & ((1 << 62) - 1)
            valsum += float(C.values.sum(dtype=torch.float64).item())
        del C
        done += 1
        if max_chunks and done >= max_chunks:
            break
    ms = c.allmax(spent * 1e3)
    per_rank_s = [round(spent, 3)]
    if c.world > 1:
        per_rank_s = [None] * c.world
        c.dist.all_gather_object(per_rank_s, round(spent, 3))
    tot = c.allsum([blocks, nnz, cand, surv, done], torch.int64)
    ks = c.allsum([keysum], torch.int64)[0] & ((1 << 62) - 1)
    vs = c.allsum([valsum])[0]
    errs = [err]
    if c.world > 1:
        errs = [None] * c.world
        c.dist.all_gather_object(errs, err)
    complete = not any(errs) and not max_chunks
    return {"value": (flops / ms / 1e6) if complete else None, "unit": "GFLOP/s incl. symbolic", "ms": ms, "flops": flops, "candidate_pairs_total": cand_total,
            "chunks_per_rank": cpr, "chunks_done": tot[4], "c_blocks": tot[0], "c_nnz": tot[1], "surviving_pairs": tot[3], "checksum_keys": ks,
            "checksum_values": vs, "errors": [e for e in errs if e], "per_rank_s": per_rank_s,
            "c_handling": "every chunk's C reduced to (blocks, values, sum of keys, sum of values) and dropped; the sums are independent of N"}


def poisson_split(c: Ctx, grid, steps):
    """P4096 cut into N slabs of grid/N grid rows (strong scaling of BASELINE config 2)"""
    from bmsparse_spgemm_spmv_b200.dist import ShardedSpMV
    B, G, torch = c.B, c.G, c.torch
    n = grid * grid
    rows_per = (grid // c.world) * grid
    lo = c.rank * rows_per
    hi = n if c.rank == c.world - 1 else lo + rows_per
    i = np.arange(lo, hi, dtype=np.int64)
    xg = i % grid; yg = i // grid
    cols = np.stack([i - grid, i - 1, i, i + 1, i + grid], axis=1)
    valid = np.stack([yg > 0, xg > 0, np.ones(i.size, bool), xg < grid - 1, yg < grid - 1], axis=1)
    vals = np.broadcast_to(np.array([-1, -1, 4, -1, -1], np.float32), (i.size, 5))
    lrp = np.zeros(i.size + 1, np.int64); np.cumsum(valid.sum(1), out=lrp[1:])
    x0 = G.x_vector(n)[lo:hi]
    nnz = int(valid.sum())
    if c.world == 1:
        A = B.bmSpMatrix.from_csr(n, n, lrp.astype(np.int32), cols[valid].astype(np.int32), vals[valid].copy())
        x = torch.from_numpy(x0).to(c.dev); y = torch.empty(n, device=c.dev)
        ms = time_steps(c, lambda: B.bmSparse_SpMV(A, x, y), steps)
        nbytes = A.spmv_bytes()
    else:
        bounds = np.arange(c.world + 1, dtype=np.int64) * rows_per
        bounds[-1] = n
        sh = ShardedSpMV(bounds, (lrp.astype(np.int32), cols[valid], vals[valid].copy()), n, device=c.dev)
        sh.set_x(torch.from_numpy(x0).to(c.dev))
        ms = time_steps(c, sh.step, steps)
        sh.check()
        sh.close()
        nbytes = 444452864 if grid == 4096 else None
    return {"ms_per_step": ms, "value": (nbytes / ms / 1e6) if nbytes else None, "unit": "GB/s (algorithmic bytes of the whole matrix)", "nnz_local": nnz}


def run(B, G, torch, dist, rank, world, dev, left, skipped, scale=22, steps=50):
    """bench.py's "strong" section.  left(): seconds of wall-clock budget remaining (rank 0 decides for everybody)."""
    c = Ctx(B, G, torch, dist, rank, world, dev)
    out = {"n_gpus": world, "note": "same matrices at every N; speed-up = this value / the N = 1 value of the same lease"}
    if c.agree(left() > 25):
        out["p4096_spmv_split"] = poisson_split(c, 4096, steps)
    else:
        skipped["strong.p4096_spmv_split"] = "time budget"
    if not c.agree(left() > 60):
        skipped["strong.rmat22"] = "time budget"
        return out
    t0 = time.perf_counter()
    n, _, rp_d, ci_d, v_d = G.rmat_torch(scale, device=dev)       # bit-identical to generators.rmat (numpy), seconds instead of minutes
    torch.cuda.synchronize()
    A = B.bmSpMatrix.from_csr(n, n, rp_d, ci_d, v_d)
    rp, ci, v = rp_d.cpu().numpy(), ci_d.cpu().numpy(), v_d.cpu().numpy()
    base = {"rows": n, "nnz": int(ci.size), "blocks": A.block_num, "generate_s": round(time.perf_counter() - t0, 1)}
    nbytes = A.spmv_bytes()
    ms, info = sharded_spmv(c, A, n, rp, ci, v, G.x_vector(n), steps, diagnostics=world > 1)
    out[f"rmat{scale}_spmv"] = dict(base, value=nbytes / ms / 1e6, unit="GB/s (algorithmic bytes of the whole matrix)", ms_per_step=ms, algorithmic_bytes=nbytes, **info)
    del v
    # A*A takes ~21 s on one GPU, ~3 s on eight
    need = 45.0 / world + 15
    if c.agree(left() > need):
        Bt = B.bmSpMatrix.from_csr(n, n, rp_d, ci_d, v_d, transpose=True)
        del rp_d, ci_d, v_d
        out[f"rmat{scale}_spgemm"] = dict(base, **rmat_spgemm(c, A, Bt, rp, ci))
    else:
        skipped[f"strong.rmat{scale}_spgemm"] = f"time budget -- needs about {need:.0f} s"
    return out
