"""Worker of tests/test_gpu_dist.py (one rank per GPU under torchrun): sharded SpMV with the halo exchange over peer memory
(fused into the kernel for the row-tiled path, wait/push kernels around the block-parallel path) against the NCCL
send/recv exchange (bit-identical: same local matrices) and against the single-GPU product of the whole matrix."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bmsparse_spgemm_spmv_b200 as B  # noqa: E402
from bmsparse_spgemm_spmv_b200.dist import ShardedSpMV, balanced_block_row_bounds, csr_row_slice, split_by_weight  # noqa: E402

G = B.generators


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    steps = 6
    cases = {"poisson": G.poisson5pt(512, 96 * world), "clustered": G.block_clustered(3000 * world), "rmat": G.rmat(13),
             "uniform": G.uniform_random(20000, 8)}
    for name, (n, _, rp, ci, v) in cases.items():
        if name == "rmat":                      # scaled down so that repeated products stay finite
            v = (v * 0.01).astype(np.float16).astype(np.float32)
        if name in ("uniform", "clustered"):
            v = (v * 0.1).astype(np.float16).astype(np.float32)
        if name == "rmat":
            # the measured-cost split of a scattered matrix: a valid partition, the same on every rank
            Aw = B.bmSpMatrix.from_csr(n, n, rp, ci, v)
            bb = balanced_block_row_bounds(Aw, iters=2, reps=3)
            assert bb[0] == 0 and bb[-1] == (n + 7) // 8 and np.all(np.diff(bb) >= 0)
            t = torch.from_numpy(np.asarray(bb, np.int64)).to(dev); t0 = t.clone()
            dist.broadcast(t0, 0)
            assert torch.equal(t, t0), "measured-cost bounds differ between ranks"
            bounds = np.asarray(bb, np.int64) * 8
            del Aw
        else:
            w = np.add.reduceat(np.diff(rp).astype(np.float64), np.arange(0, n, 8)) + 1.0
            bounds = split_by_weight(w, world) * 8
        bounds[-1] = n
        lcsr = csr_row_slice(rp, ci, v, int(bounds[rank]), int(bounds[rank + 1]))
        x0 = G.x_vector(n)
        res = {}
        for mode in ("p2p", "nccl"):
            sh = ShardedSpMV(bounds, lcsr, n, device=dev, halo=mode)
            assert (sh.p2p is not None) == (mode == "p2p")
            sh.set_x(torch.from_numpy(x0[bounds[rank]:bounds[rank + 1]]).to(dev))
            for _ in range(steps):
                sh.step()
            torch.cuda.synchronize()
            sh.check()
            res[mode] = sh.y_own().clone()
            sh.close()
        assert torch.equal(res["p2p"], res["nccl"]), f"{name}: peer-memory and NCCL exchanges differ on rank {rank}"
        # whole matrix on this GPU
        A = B.bmSpMatrix.from_csr(n, n, rp, ci, v)
        x = torch.from_numpy(x0).to(dev)
        for _ in range(steps):
            x = B.bmSparse_SpMV(A, x)
        ref = x[bounds[rank]:bounds[rank + 1]]
        scale = float(x.abs().max()) + 1e-30
        err = float((res["p2p"] - ref).abs().max()) / scale
        assert np.isfinite(scale) and err < 1e-4, f"{name}: rank {rank} rel err {err} (scale {scale})"
        if rank == 0:
            print(f"DIST_OK {name} world={world} rows={n} err={err:.2e}", flush=True)
    one_directional(rank, world, dev)
    dist.barrier()
    dist.destroy_process_group()


def one_directional(rank, world, dev):
    """lower-bidiagonal matrix -- rank r pushes its last row to rank r+1 and reads nothing remote, so only the epoch
    wait of the pushing tile keeps it from running products ahead of a slow reader and overwriting a buffer still in use.
    Many steps, the last rank deliberately slowed; the result must equal the single-GPU iteration."""
    n = 4096 * world
    i = np.arange(n, dtype=np.int64)
    cols = np.stack([i - 1, i], axis=1); valid = np.stack([i > 0, np.ones(n, bool)], axis=1)
    ci = cols[valid].astype(np.int32); v = np.full(ci.size, 0.5, np.float32)
    rp = np.zeros(n + 1, np.int64); np.cumsum(valid.sum(1), out=rp[1:])
    bounds = np.arange(world + 1, dtype=np.int64) * 4096
    lcsr = csr_row_slice(rp.astype(np.int32), ci, v, int(bounds[rank]), int(bounds[rank + 1]))
    x0 = G.x_vector(n)
    steps = 120
    sh = ShardedSpMV(bounds, lcsr, n, device=dev, halo="p2p")
    assert sh.local.nnz / sh.local.block_num >= 2.5            # row-tiled path: the fused kernel
    sh.set_x(torch.from_numpy(x0[bounds[rank]:bounds[rank + 1]]).to(dev))
    for _ in range(steps):
        if rank == world - 1:
            torch.cuda._sleep(400000)                          # ~0.2 ms: the reader falls behind
        sh.step()
    torch.cuda.synchronize()
    sh.check()
    got = sh.y_own().clone()
    sh.close()
    A = B.bmSpMatrix.from_csr(n, n, rp.astype(np.int32), ci, v)
    x = torch.from_numpy(x0).to(dev)
    for _ in range(steps):
        x = B.bmSparse_SpMV(A, x)
    assert torch.equal(got, x[bounds[rank]:bounds[rank + 1]]), f"one-directional: rank {rank} differs from the single-GPU iteration"
    if rank == 0:
        print(f"DIST_OK one_directional world={world} steps={steps}", flush=True)


if __name__ == "__main__":
    main()
