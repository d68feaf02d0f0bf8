#!/usr/bin/env python
"""BASELINE config 5: R-MAT (a,b,c,d = .57,.19,.19,.05; 16 edges/vertex, duplicates merged) SpMV and A*A SpGEMM,
block-row sharded over the GPUs of one box (strong scaling: the same matrix at every N).  Command-line front end of
tools/strong_scaling.py (bench.py's "strong" section runs the same functions).

  python tools/rmat_scale.py --scale 22                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 ... tools/rmat_scale.py --scale 22

Prints one JSON line per operator (rank 0)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bmsparse_spgemm_spmv_b200 as B  # noqa: E402
from tools import strong_scaling as S  # noqa: E402

G = B.generators


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--what", default="spmv,spgemm,poisson")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--chunk-pairs", type=float, default=3e9, help="candidate pairs per bmsp_spgemm call: large enough that a chunk holds many block rows next to its hub rows, small enough that C stays below 2^31 blocks / 2^32 values and fits HBM")
    ap.add_argument("--max-chunks", type=int, default=0, help="stop the SpGEMM after this many chunks per rank (0 = all): bounded sample")
    ap.add_argument("--halo", default="auto")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = S.Ctx(B, G, torch, dist if world > 1 else None, rank, world, dev)
    emit = lambda d: rank == 0 and print(json.dumps(d), flush=True)
    if "poisson" in a.what:
        emit(dict(op="p4096_spmv_split", n_gpus=world, **S.poisson_split(c, 4096, a.steps)))
    if "spmv" in a.what or "spgemm" in a.what:
        t0 = time.perf_counter()
        n, _, rp_d, ci_d, v_d = G.rmat_torch(a.scale, device=dev)
        torch.cuda.synchronize()
        A = B.bmSpMatrix.from_csr(n, n, rp_d, ci_d, v_d)
        rp, ci, v = rp_d.cpu().numpy(), ci_d.cpu().numpy(), v_d.cpu().numpy()
        base = {"rows": n, "nnz": int(ci.size), "blocks": A.block_num, "n_gpus": world, "scale": a.scale, "generate_s": round(time.perf_counter() - t0, 1)}
        if "spmv" in a.what:
            nbytes = A.spmv_bytes()
            ms, info = S.sharded_spmv(c, A, n, rp, ci, v, G.x_vector(n), a.steps, halo=a.halo, diagnostics=world > 1)
            emit(dict(base, op="spmv", metric="SpMV HBM GB/s", value=nbytes / ms / 1e6, ms_per_step=ms, algorithmic_bytes=nbytes, scaling="strong", **info))
        if "spgemm" in a.what:
            Bt = B.bmSpMatrix.from_csr(n, n, rp_d, ci_d, v_d, transpose=True)
            del rp_d, ci_d, v_d
            emit(dict(base, op="spgemm", metric="SpGEMM GFLOP/s incl. symbolic", scaling="strong", **S.rmat_spgemm(c, A, Bt, rp, ci, a.chunk_pairs, a.max_chunks)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
