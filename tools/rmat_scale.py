#!/usr/bin/env python
"""BASELINE config 5: R-MAT (a,b,c,d = .57,.19,.19,.05; 16 edges/vertex, duplicates merged) SpMV and A*A SpGEMM,
block-row sharded over the GPUs of one box (strong scaling: the same matrix at every N).

  python tools/rmat_scale.py --scale 22                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 ... tools/rmat_scale.py --scale 22

SpMV   rows split by SpMV bytes (bmsp_partition_block_rows); every rank keeps x for its extended column range, the
       peers' slices arrive over NVLink peer memory (bmsp_spmv_halo; for R-MAT every rank needs every slice, so
       the exchange is an all-gather written by the producers).  value = algorithmic bytes of the WHOLE matrix /
       max-over-ranks device time per product.
SpGEMM A's block rows split by candidate pairs into N x k chunks dealt to the ranks block-cyclically; B^t replicated; chunks of
       <= --chunk-pairs candidate pairs (default 3e9; the scale-22 product, ~7e10 values, fits no GPU: each chunk's C is reduced
       to a checksum -- blocks, values, sum of keys, sum of values -- and dropped).  The checksums are summed over
       ranks and are independent of N.  value = 2 * scalar products / max-over-ranks time of the bmsp_spgemm calls.
Prints one JSON line per operator (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bmsparse_spgemm_spmv_b200 as B  # noqa: E402
from bmsparse_spgemm_spmv_b200.dist import ShardedSpMV, csr_row_slice  # noqa: E402

G = B.generators


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--what", default="spmv,spgemm")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--chunk-pairs", type=float, default=3e9, help="candidate pairs per bmsp_spgemm call: large enough that a chunk holds many block rows next to its hub rows (3e8: RM20 11.6 s; one call for all of RM18 is 4x faster than 12 chunks), small enough that C stays below 2^31 blocks / 2^32 values and fits HBM")
    ap.add_argument("--max-chunks", type=int, default=0, help="stop the SpGEMM after this many chunks per rank (0 = all): bounded sample")
    ap.add_argument("--halo", default="auto")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = sys.stdout

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def allsum(vs, dtype=torch.float64):
        t = torch.tensor(vs, dtype=dtype, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().tolist()

    t0 = time.perf_counter()
    n, _, rp_d, ci_d, v_d = G.rmat_torch(a.scale, device=dev)       # bit-identical to generators.rmat (numpy), seconds instead of minutes
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    d = lambda x: torch.from_numpy(x).to(dev)
    A = B.bmSpMatrix.from_csr(n, n, rp_d, ci_d, v_d)
    rp, ci, v = rp_d.cpu().numpy(), ci_d.cpu().numpy(), v_d.cpu().numpy()
    nbr = A.num_block_rows
    base = {"rows": n, "nnz": int(ci.size), "blocks": A.block_num, "n_gpus": world, "scale": a.scale, "generate_s": round(gen_s, 1)}

    if "spmv" in a.what:
        nbytes = A.spmv_bytes()
        if world == 1:
            x = d(G.x_vector(n)); y = torch.empty(n, device=dev)
            step = lambda: B.bmSparse_SpMV(A, x, y)
            sh = None
        else:
            bounds = A.partition_block_rows(world).astype(np.int64) * 8
            bounds[-1] = n
            lcsr = csr_row_slice(rp, ci, v, int(bounds[rank]), int(bounds[rank + 1]))
            sh = ShardedSpMV(bounds, lcsr, n, device=dev, halo=a.halo)
            sh.set_x(d(G.x_vector(n)[bounds[rank]:bounds[rank + 1]]))
            step = sh.step
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record(); e1.synchronize()
        ms = allmax(e0.elapsed_time(e1) / a.steps)
        local = None
        if sh is not None:
            sh.check()
            # per-rank diagnostics: the local product alone (no exchange), rows and blocks of the shard
            xl = sh.x[sh.cur]; yl = torch.empty(sh.own_hi - sh.own_lo, device=dev)
            for _ in range(3):
                B.bmSparse_SpMV(sh.local, xl, yl)
            e0.record()
            for _ in range(a.steps):
                B.bmSparse_SpMV(sh.local, xl, yl)
            e1.record(); e1.synchronize()
            mine = {"rank": rank, "local_ms": round(e0.elapsed_time(e1) / a.steps, 4), "rows": sh.own_hi - sh.own_lo, "blocks": sh.local.block_num}
            local = [None] * world
            dist.all_gather_object(local, mine)
        line = dict(base, op="spmv", per_rank=local, metric="SpMV HBM GB/s", value=nbytes / ms / 1e6, ms_per_step=ms, algorithmic_bytes=nbytes,
                    halo=("peer-memory" if sh is not None and sh.p2p is not None else ("nccl" if sh is not None else "none")),
                    halo_bytes_in_per_rank=(sh.halo_bytes if sh is not None else 0), scaling="strong")
        if rank == 0:
            print(json.dumps(line), file=out, flush=True)
        if sh is not None:
            sh.close()

    if "spgemm" in a.what:
        Bt = B.bmSpMatrix.from_csr(n, n, rp_d, ci_d, v_d, transpose=True)
        del rp_d, ci_d, v_d
        rowlen = np.diff(rp).astype(np.int64)
        flops = 2 * int(rowlen[ci].sum())
        # total candidate pairs -> chunk count; the same partition call splits ranks and chunks
        brp = A.block_row_ptr.cpu().numpy().astype(np.int64); bcol = A.block_col.cpu().numpy()
        blen = np.diff(Bt.block_row_ptr.cpu().numpy().astype(np.int64))
        cand_total = int(blen[bcol].sum())
        cpr = max(1, int(np.ceil(cand_total / world / a.chunk_pairs)))
        bounds = A.partition_block_rows(world * cpr, Bt)
        # block-cyclic: rank r multiplies chunks r, r + N, r + 2N, ... -- the hub rows sit in the first chunks, and a hub chunk costs
        # more per candidate pair than a tail chunk, so contiguous ranges would leave rank 0 with the slowest ones
        mine = range(rank, world * cpr, world)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        spent = 0.0; blocks = 0; nnz = 0; keysum = 0; valsum = 0.0; cand = 0; surv = 0; done = 0; err = None; worst = (0.0, 0, 0)
        for c in mine:
            r0, r1 = int(bounds[c]), int(bounds[c + 1])
            if r1 <= r0:
                continue
            torch.cuda.synchronize(); t = time.perf_counter()
            try:
                C, info = B.bmSparse_mult(A, Bt, None, 0, False, 5, brow_range=(r0, r1))
            except B.BmspError as e:
                err = f"rows [{r0},{r1}): {e}"
                break
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            spent += dt
            if dt > worst[0]:
                worst = (dt, r0, r1)
            blocks += C.block_num; nnz += C.nnz; cand += info.candidate_pairs; surv += info.surviving_pairs
            if C.block_num:
                keysum = (keysum + int(C.keys.sum().item())) & ((1 << 62) - 1)
                valsum += float(C.values.sum(dtype=torch.float64).item())
            del C
            done += 1
            if a.max_chunks and done >= a.max_chunks:
                break
        ms = allmax(spent * 1e3)
        tot = allsum([blocks, nnz, cand, surv, done], torch.int64)
        ks = allsum([keysum], torch.int64)[0] & ((1 << 62) - 1)
        vs = allsum([valsum])[0]
        errs = [None] * world
        if world > 1:
            dist.all_gather_object(errs, err)
        else:
            errs = [err]
        complete = not any(errs) and not a.max_chunks
        line = dict(base, op="spgemm", metric="SpGEMM GFLOP/s incl. symbolic", value=(flops / ms / 1e6) if complete else None, ms=ms, flops=flops,
                    candidate_pairs_total=cand_total, chunks_per_rank=cpr, chunks_done=tot[4], c_blocks=tot[0], c_nnz=tot[1],
                    candidate_pairs=tot[2], surviving_pairs=tot[3], checksum_keys=ks, checksum_values=vs, scaling="strong",
                    slowest_chunk_rank0={"s": round(worst[0], 3), "rows": [worst[1], worst[2]]}, errors=[e for e in errs if e],
                    c_handling="every chunk's C reduced to (blocks, values, sum of keys, sum of values) and dropped")
        if rank == 0:
            print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
