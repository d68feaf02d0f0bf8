#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the bmSparse hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline (BASELINE.json configs[1]): bmSparse SpMV, fp16 values / fp32 x, y / fp32 accumulate, synthetic
2-D Poisson 4096x4096 grid (16.7M rows, 83.9M nnz) -- metric "SpMV HBM GB/s" = algorithmic bytes of the
compact surface (nblk*12 + nbr*8 + nnz*2 + ncols*4 + nrows*4 = 444 452 864 B) / device time.
A step = one SpMV.  N > 1 (torchrun, one rank per GPU): weak scaling, every rank owns one 4096x4096-grid
slab of a 4096 x (4096 N) grid; the SpMV kernel itself pushes its boundary rows of y into the neighbours' next x over
NVLink peer memory (bmsp_spmv_halo; NCCL send/recv when peer mapping is unavailable or BMSP_BENCH_HALO=nccl).
The same JSON line also carries the SpGEMM result (BASELINE configs[2], uniform-random 1M x 1M, 16 nnz/row,
A*A, GFLOP/s including the symbolic phase) under "spgemm" -- measured on rank 0 at N = 1 only.

--impl reference times the reference's own CPU path (cusp host CSR kernels compiled from the reference tree
into oracle/_ref; the oracle port if that library is missing) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 4096
_OUT = sys.stdout
SPGEMM_N, SPGEMM_K = 1_000_000, 16


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """polls NVML (clocks + event reasons) while the timed region runs"""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples = []   # (t, sm_mhz, reasons)
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(0.0005)

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        where = "timed region"
        if not inside:
            inside = [s for s in self.samples if s[0] <= t1][-5:]
            where = "last samples before the timed region ended (region shorter than the NVML poll)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "where": "nvml unavailable"}
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [n for b, n in self.REASONS.items() if mask & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "where": where}


def cpu_spmv_baseline(O, nr, nc, rp, ci, v, x, nbytes, reps):
    """reference cusp host CSR SpMV (sequential = thrust::cpp dispatch, and the OMP back-end) on this box."""
    have_ref = O.ref_lib() is not None
    best = None
    threads = O.max_threads()
    for omp in (False, True):
        ts = []
        for i in range(reps + 1):
            t = time.perf_counter()
            if have_ref:
                O.ref_csr_spmv(nr, nc, rp, ci, v, x, omp=omp)
            else:
                O.csr_spmv(rp, ci, v, x, threads=threads if omp else 1)
            if i:
                ts.append(time.perf_counter() - t)
        t = float(np.median(ts))
        if best is None or t < best[0]:
            best = (t, threads if omp else 1)
    return {"value": nbytes / best[0] / 1e9, "unit": "GB/s", "cores": best[1], "kind": "reference" if have_ref else "port",
            "seconds_per_step": best[0], "host_threads_available": threads,
            "sample": f"full workload: {reps} CSR SpMVs of the {nr}-row matrix, fp32, cusp sequential and OMP back-ends, best kept; "
                      f"numerator = the metric's {nbytes} algorithmic bytes"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from bmsparse_spgemm_spmv_b200 import generators as G
    nr, nc, rp, ci, v = G.poisson5pt(GRID, GRID)
    x = G.x_vector(nc)
    nblk_bytes = 444452864
    steps = max(1, min(args.steps, 20))
    for _ in range(max(args.warmup, 1)):
        O.ref_csr_spmv(nr, nc, rp, ci, v, x) if O.ref_lib() is not None else O.csr_spmv(rp, ci, v, x)
    base = cpu_spmv_baseline(O, nr, nc, rp, ci, v, x, nblk_bytes, steps)
    line = {"impl": "reference", "metric": "SpMV HBM GB/s", "value": base["value"], "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": base["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"poisson5pt {GRID}x{GRID} CSR SpMV on host memory (cusp::multiply host path), one slab per step"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


def spgemm_bench(B, G, torch, O, steps=3):
    nr, nc, rp, ci, v = G.uniform_random(SPGEMM_N, SPGEMM_K, seed=2)
    d = lambda a: torch.from_numpy(a).cuda()
    t0 = time.perf_counter()
    A = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v))
    Bt = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v), transpose=True)
    torch.cuda.synchronize()
    conv_ms = (time.perf_counter() - t0) * 1e3
    rowlen = np.diff(rp).astype(np.int64)
    flops = 2 * int(rowlen[ci].sum())
    times, info = [], None
    for i in range(steps + 1):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        C, info = B.bmSparse_mult(A, Bt, None, 0, True, 5)
        e1.record(); e1.synchronize()
        if i:
            times.append(e0.elapsed_time(e1))
        cb, cn = C.block_num, C.nnz
        del C
    ms = float(np.median(times))
    out = {"workload": f"uniform-random {SPGEMM_N}x{SPGEMM_N}, {SPGEMM_K} nnz/row, A*A, fp16 in / fp32 out", "metric": "SpGEMM GFLOP/s incl. symbolic",
           "gflops": flops / ms / 1e6, "ms": ms, "symbolic_ms": info.symbolic_ms, "numeric_ms": info.numeric_ms, "flops": flops,
           "candidate_pairs": info.candidate_pairs, "surviving_pairs": info.surviving_pairs, "c_blocks": cb, "c_nnz": cn,
           "convert_ms_both_operands": conv_ms}
    # CPU baseline on a bounded sample: the first 1/16 of A's rows times the full B
    if O is not None:
        rows = nr // 16
        srp = rp[:rows + 1]; sci = ci[:srp[-1]]; sv = v[:srp[-1]]
        t = time.perf_counter()
        if O.ref_lib() is not None:
            O.ref_csr_spgemm(rows, nc, srp, sci, sv, nr, nc, rp, ci, v, omp=False, copy=False); kind = "reference"
        else:
            O.csr_spgemm(rows, nc, srp, sci, sv, rp, ci, v); kind = "port"
        dt = time.perf_counter() - t
        sflops = 2 * int(rowlen[sci].sum())
        out["cpu_baseline"] = {"value": sflops / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                               "sample": f"first {rows} rows of A times the full B through cusp sequential csr_spgemm ({dt:.2f} s)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-spgemm", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # native libraries (NCCL's version banner) print to fd 1: keep the real stdout for the one JSON line only
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import bmsparse_spgemm_spmv_b200 as B
    from bmsparse_spgemm_spmv_b200 import generators as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, args.warmup
    peak, peak_src = peaks()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- build the workload: one 4096 x 4096-grid slab per rank
    nr, nc, rp, ci, v = G.poisson5pt(GRID, GRID)      # slab template (interior coupling added below for N > 1)
    x_host = G.x_vector(nc, seed=1 + rank)
    sharded = None
    if world == 1:
        t0 = time.perf_counter()
        A = B.bmSpMatrix.from_csr(nr, nc, torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(v).to(dev))
        torch.cuda.synchronize()
        conv_ms = (time.perf_counter() - t0) * 1e3
        x = torch.from_numpy(x_host).to(dev)
        y = torch.empty(nr, dtype=torch.float32, device=dev)
        step = lambda: B.bmSparse_SpMV(A, x, y)
        nbytes = A.spmv_bytes()
        launches_per_step = 1
    else:
        from bmsparse_spgemm_spmv_b200.dist import ShardedSpMV
        # global grid 4096 x (4096*world): this rank's rows, with the +-4096 couplings into the neighbour slabs
        n_glob = nr * world
        base = rank * nr
        i = np.arange(nr, dtype=np.int64) + base
        xg = i % GRID; yg = i // GRID
        cols = np.stack([i - GRID, i - 1, i, i + 1, i + GRID], axis=1)
        valid = np.stack([yg > 0, xg > 0, np.ones(nr, bool), xg < GRID - 1, yg < GRID * world - 1], axis=1)
        vals = np.broadcast_to(np.array([-1, -1, 4, -1, -1], np.float32), (nr, 5))
        lrp = np.zeros(nr + 1, np.int64); np.cumsum(valid.sum(1), out=lrp[1:])
        bounds = np.arange(world + 1, dtype=np.int64) * nr
        sharded = ShardedSpMV(bounds, (lrp.astype(np.int32), cols[valid], vals[valid].copy()), n_glob, device=dev,
                              halo=os.environ.get("BMSP_BENCH_HALO", "auto"))
        sharded.set_x(torch.from_numpy(x_host).to(dev))
        A = sharded.local
        conv_ms = None
        step = sharded.step
        nbytes = A.spmv_bytes()
        launches_per_step = 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / K
    clocks = sampler.summary(t_start, t_end)

    # ---- kernel-only duration of the dominant kernel (same stream, CUDA events, x resident): N = 1 uses the step itself
    kern_ms = ms_step
    if world > 1:
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        xb, yb = sharded.x[sharded.cur], sharded.own_slice(sharded.x[(sharded.cur + 1) % len(sharded.x)])
        for _ in range(3):
            B.bmSparse_SpMV(A, xb, yb)
        ev0.record()
        for _ in range(K):
            B.bmSparse_SpMV(A, xb, yb)
        ev1.record(); ev1.synchronize()
        kern_ms = ev0.elapsed_time(ev1) / K

    # ---- end to end through the public API with HOST buffers: H2D x, SpMV, D2H y inside the timed region
    ncols_local = nc if world == 1 else A.num_cols
    xp = torch.zeros(ncols_local, dtype=torch.float32).pin_memory()
    xoff = 0 if world == 1 else sharded.own_lo - sharded.ext_lo
    xp[xoff:xoff + nr].copy_(torch.from_numpy(x_host))
    yp = torch.empty(nr, dtype=torch.float32).pin_memory()
    Ke = max(3, min(K, 50))

    def e2e_step():
        B.bmSparse_SpMV_host(A, xp, yp)               # H2D of x, the product, D2H of y: pipelined inside the library
        torch.cuda.current_stream().synchronize()     # the caller needs y on the host before the next step

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / Ke
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    sampler.stop_flag = True

    total_bytes = nbytes * world
    value = total_bytes / (ms_step * 1e-3) / 1e9
    line = {
        "metric": "SpMV HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 values, f32 x/y/accumulate",
        "data": "synthetic",
        "config": {"workload": f"bmSparse SpMV, 2-D Poisson 5-point {GRID}x{GRID} grid per GPU ({nr} rows, {A.nnz} nnz, {A.block_num} 8x8 blocks)"
                               + ("" if world == 1 else f"; global grid {GRID}x{GRID * world}, x halo ({sharded.halo_bytes} B in per rank and step) "
                                                             + ("pushed over NVLink peer memory by the SpMV kernel itself (bmsp_spmv_halo)" if sharded.p2p is not None
                                                                else "exchanged with NCCL send/recv before each product")),
                   "algorithmic_bytes_per_step_per_gpu": nbytes,
                   "l2": "inputs larger than L2 (444 MB streamed per step vs 126 MB L2); no flush between steps",
                   "timing": "CUDA events on the launching stream around K steps, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": nbytes / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": nbytes / (kern_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": nbytes / (kern_ms * 1e-3) / 1e9 / 8000.0,
                     "kernel": "spmv_tile_kernel<__half,float,64,2,12>", "kernel_ms": kern_ms},
        "e2e": {"value": total_bytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": ncols_local * 4, "d2h_bytes_per_step": nr * 4,
                "ms_per_step": e2e_s * 1e3, "note": "bmsp_spmv_host: matrix resident in HBM (as in the reference's timed region); every step x comes from pinned host memory in chunks on a copy stream while the row-range launches of the same kernel store y straight into the pinned host buffer (PCIe-bound both ways)"},
        "gpu_launches": K * launches_per_step,
        "clocks": clocks,
        "convert_ms": conv_ms,
    }
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            line["roofline"]["traffic"] = json.load(open(tf)).get("spmv_rowtile_p4096_dram_bytes")
        except Exception:
            pass

    if rank == 0 and world == 1:
        O = None
        if not args.no_cpu:
            from oracle import oracle as O
            line["cpu_baseline"] = cpu_spmv_baseline(O, nr, nc, rp, ci, v, x_host, nbytes, 5)
        if not args.no_spgemm:
            del A
            torch.cuda.empty_cache()
            line["spgemm"] = spgemm_bench(B, G, torch, O)
    if rank == 0:
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        sharded.check()
        sharded.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
