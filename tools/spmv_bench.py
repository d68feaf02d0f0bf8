"""Times bmsp_spmv on one generator config (CUDA events, inputs larger than L2).  BMSP_SPMV_VARIANT selects the kernel build.
usage: python tools/spmv_bench.py [p4096|p2048|bc|u1m|rmat20] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bmsparse_spgemm_spmv_b200 as B  # noqa: E402

G = B.generators
which = sys.argv[1] if len(sys.argv) > 1 else "p4096"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
gen = {"p4096": lambda: G.poisson5pt(4096, 4096), "p2048": lambda: G.poisson5pt(2048, 2048), "bc": lambda: G.block_clustered(1 << 19),
       "u1m": lambda: G.uniform_random(1_000_000, 16), "rmat20": lambda: G.rmat(20), "rmat22": lambda: G.rmat(22)}[which]
nr, nc, rp, ci, v = gen()
d = lambda a: torch.from_numpy(a).cuda()
A = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v))
x = d(G.x_vector(nc))
y = torch.empty(nr, dtype=torch.float32, device="cuda")
t0 = time.perf_counter()
B.bmSparse_SpMV(A, x, y)
torch.cuda.synchronize()
plan_ms = (time.perf_counter() - t0) * 1e3
for _ in range(10):
    B.bmSparse_SpMV(A, x, y)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    B.bmSparse_SpMV(A, x, y)
e1.record(); e1.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
nbytes = A.spmv_bytes()
print(f"SPMV {which} variant={os.environ.get('BMSP_SPMV_VARIANT', '0')} rows={nr} nnz={A.nnz} blocks={A.block_num} "
      f"us={us:.2f} GBps={nbytes / us / 1e3:.1f} frac6468={nbytes / us / 1e3 / 6468.3:.3f} first_call_ms={plan_ms:.1f} ysum={float(y.double().sum()):.6e}")
