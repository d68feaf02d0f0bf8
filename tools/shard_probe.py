"""One-GPU look at the shards an N-way R-MAT-22 SpMV run gives its ranks: partition the whole matrix N ways (the cost model of
bmsp_partition_block_rows), slice every shard and time its local product alone.  usage: python tools/shard_probe.py [N] [steps] [scale]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bmsparse_spgemm_spmv_b200 as B  # noqa: E402

G = B.generators
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
scale = int(sys.argv[3]) if len(sys.argv) > 3 else 22
n, _, rp, ci, v = G.rmat_torch(scale)
A = B.bmSpMatrix.from_csr(n, n, rp, ci, v)
x = torch.from_numpy(G.x_vector(n)).cuda()
bounds = A.partition_block_rows(N)
print("bounds (block rows)", bounds.tolist())


def timeit(fn):
    for _ in range(5):
        fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


y = torch.empty(n, device="cuda")
print(f"whole matrix: {timeit(lambda: B.bmSparse_SpMV(A, x, y)):.1f} us  blocks {A.block_num}")
for p in range(N):
    S = A.slice_block_rows(int(bounds[p]), int(bounds[p + 1]), rebase=True)
    ys = torch.empty(S.num_rows, device="cuda")
    us = timeit(lambda: B.bmSparse_SpMV(S, x, ys))
    print(f"shard {p}: rows {S.num_rows} blocks {S.block_num} local {us:.1f} us")
    del S
