"""CPU suite: the N>1 host logic (row split, range-halo plan, ping-pong exchange) over gloo, world_size 2 and 3.
The local product is injected (the oracle's cusp CSR SpMV) -- this tests the sharding plumbing, not the kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bmsparse_spgemm_spmv_b200 import generators as G
        from bmsparse_spgemm_spmv_b200.dist import ShardedSpMV, csr_row_slice, split_by_weight
        from oracle import oracle as O
        if kind == "poisson":
            n, _, rp, ci, v = G.poisson5pt(40, 36)
        else:
            n, _, rp, ci, v = G.rmat(9)
        w = np.add.reduceat(np.diff(rp).astype(np.float64), np.arange(0, n, 8)) + 1.0
        bounds = split_by_weight(w, world) * 8
        bounds[-1] = n
        lrp, lci, lv = csr_row_slice(rp, ci, v, int(bounds[rank]), int(bounds[rank + 1]))

        def build(nr, nc, rp_, ci_, v_):
            return (rp_, ci_, v_)

        holder = {}

        def spmv(x_ext, y_out):
            rp_, ci_, v_ = holder["m"]
            y_out.copy_(torch.from_numpy(O.csr_spmv(rp_, ci_, v_, x_ext.numpy())))

        sh = ShardedSpMV(bounds, (lrp, lci, lv), n, spmv_fn=spmv, build_fn=build)
        holder["m"] = sh.local
        x0 = G.x_vector(n)
        sh.set_x(torch.from_numpy(x0[bounds[rank]:bounds[rank + 1]]))
        ref = x0.copy()
        for _ in range(3):
            sh.step()
            ref = O.csr_spmv(rp, ci, v, ref)
        got = sh.y_own().numpy()
        ok = np.array_equal(got, ref[bounds[rank]:bounds[rank + 1]])
        banded_ok = True
        if kind == "poisson" and world > 2:
            banded_ok = len(sh.recv) <= 2 and sh.halo_bytes <= 2 * 48 * 4   # neighbours only, one grid line (+align) each
        q.put((rank, bool(ok), bool(banded_ok), sh.halo_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "poisson"), (3, "poisson"), (2, "rmat"), (3, "rmat")])
def test_sharded_spmv_gloo(world, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert all(b for _, _, b, _ in res), res


def test_split_by_weight():
    from bmsparse_spgemm_spmv_b200.dist import split_by_weight
    w = np.ones(100); w[:10] = 50
    b = split_by_weight(w, 4)
    assert b[0] == 0 and b[-1] == 100 and np.all(np.diff(b) >= 0)
    loads = [w[b[i]:b[i + 1]].sum() for i in range(4)]
    assert max(loads) <= 1.4 * sum(loads) / 4
    assert np.all(split_by_weight(np.ones(64), 8, align=8) % 8 == 0)
