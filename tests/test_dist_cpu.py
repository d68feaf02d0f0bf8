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


def test_deal_chunks_partitions_every_chunk_once():
    """work chunks of the chunked SpGEMM: every chunk to exactly one rank, counts within one of each other, and the sum of a
    decreasing cost sequence is better balanced than with a plain cyclic deal"""
    from bmsparse_spgemm_spmv_b200.dist import deal_chunks
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 21, 84, 88):
            got = [deal_chunks(n, world, r) for r in range(world)]
            flat = sorted(c for g in got for c in g)
            assert flat == list(range(n))
            sizes = [len(g) for g in got]
            assert max(sizes) - min(sizes) <= 1
    cost = np.linspace(2.0, 1.0, 84)
    zig = [cost[deal_chunks(84, 4, r)].sum() for r in range(4)]
    cyc = [cost[list(range(r, 84, 4))].sum() for r in range(4)]
    assert max(zig) - min(zig) < max(cyc) - min(cyc)


def test_halo_descriptor_layout():
    """The peer-memory halo plan (pure address arithmetic of dist.halo_descriptor): three slabs of a banded matrix.  Every push
    range starts on a multiple of 4 rows, lands 16-byte aligned inside the peer's x buffer at the row's position in the peer's
    extended range, and the epoch slots are mutual (my slot in the peer's inbox / the peer's slot in mine)."""
    from bmsparse_spgemm_spmv_b200.dist import halo_descriptor
    from bmsparse_spgemm_spmv_b200 import _lib as L
    n, band = 3 * 4096, 64
    bounds = [0, 4096, 8192, n]
    ext = [(max(0, bounds[r] - band), min(n, bounds[r + 1] + band)) for r in range(3)]
    base = {0: 0x10000000, 1: 0x20000000, 2: 0x30000000}
    stride = {r: ((ext[r][1] - ext[r][0]) * 4 + 255) // 256 * 256 for r in range(3)}
    descs = {}
    for r in range(3):
        own_lo, own_hi = bounds[r], bounds[r + 1]
        send = []
        for p in range(3):
            if p != r:
                a, e = max(ext[p][0], own_lo), min(ext[p][1], own_hi)
                if a < e:
                    send.append((p, a, e))
        peers = sorted({p for p, _, _ in send})
        for buf in range(3):
            d = halo_descriptor(buf, r, send, peers, own_lo, own_hi, ext[r][0], base[r], base,
                                {p: (stride[p], ext[p][0]) for p in peers}, 0x999000)
            descs[(r, buf)] = (d, send, peers)
            assert d.n_push == len(send) <= L.HALO_MAX and d.n_peer == len(peers)
            assert d.own_col_lo == own_lo - ext[r][0] and d.own_col_hi == own_hi - ext[r][0]
            for i, (p, a, e) in enumerate(send):
                assert d.push_lo[i] % 4 == 0 and 0 <= d.push_lo[i] < d.push_hi[i] <= own_hi - own_lo
                assert d.push_dst[i] % 16 == 0
                # the destination is row `a` inside p's buffer `buf`
                assert d.push_dst[i] == base[p] + 1024 + buf * stride[p] + (a - ext[p][0]) * 4
                assert d.push_dst[i] + (e - a) * 4 <= base[p] + 1024 + (buf + 1) * stride[p]
    # middle rank talks to both neighbours, the outer ranks to one; slots are mutual
    assert descs[(1, 0)][2] == [0, 2] and descs[(0, 0)][2] == [1] and descs[(2, 0)][2] == [1]
    d0, d1 = descs[(0, 1)][0], descs[(1, 1)][0]
    assert d0.peer_flag[0] == base[1] + 4 * 0 and d1.my_flag[0] == base[1] + 4 * 0
    assert d1.peer_flag[0] == base[0] + 4 * 1 and d0.my_flag[0] == base[0] + 4 * 1
