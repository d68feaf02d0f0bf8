#!/usr/bin/env python
"""SpGEMM A*A timing on one of the synthetic configs (u1m | bc4m | bc512k | p1024 | p4096 | rmat18 ...)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bmsparse_spgemm_spmv_b200 as B
G = B.generators

def make(name):
    if name == "u1m": return G.uniform_random(1_000_000, 16, seed=2)
    if name == "u100k": return G.uniform_random(100_000, 16, seed=2)
    if name == "bc4m": return G.block_clustered(524288)
    if name == "bc512k": return G.block_clustered(65536)
    if name.startswith("p"): m = int(name[1:]); return G.poisson5pt(m, m)
    if name.startswith("rmat"): return G.rmat(int(name[4:]))
    raise SystemExit(name)

ap = argparse.ArgumentParser(); ap.add_argument("name"); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--path", type=int, default=-1)
a = ap.parse_args()
nr, nc, rp, ci, v = make(a.name)
d = lambda x: torch.from_numpy(x).cuda()
A = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v)); Bt = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v), transpose=True)
rowlen = np.diff(rp).astype(np.int64); flops = 2 * int(rowlen[ci].sum())
print(f"{a.name}: rows {nr} nnz {ci.size} blocks {A.block_num} nnz/blk {ci.size / max(A.block_num,1):.2f} flops {flops/1e6:.1f}M")
for i in range(a.reps + 1):
    torch.cuda.synchronize(); t = time.perf_counter()
    C, info = B.bmSparse_mult(A, Bt, None, 0, True, 5, numeric_path=a.path)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
    print(f"  rep {i}: wall {dt:.2f} ms  symbolic {info.symbolic_ms:.2f} numeric {info.numeric_ms:.2f}  -> {flops / dt / 1e6:.2f} GFLOP/s  cand {info.candidate_pairs} surv {info.surviving_pairs} Cblk {info.c_blocks} Cnnz {info.c_nnz} path {info.numeric_path}")
    del C
