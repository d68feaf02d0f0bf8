#!/usr/bin/env python
"""Same-box comparison: this repo's kernels vs the reference's own bmSparse CUDA operators (rebuilt for sm_100 in
oracle/_ref) vs cuSPARSE 12.x, on the BASELINE.json configs.  Writes gpurun_out/compare_<tag>.json.
Test/benchmark infrastructure: may use oracle/ (the reference runner lives there)."""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bmsparse_spgemm_spmv_b200 as B
from oracle import oracle as O
G = B.generators
CUSP = os.path.join(ROOT, "tools", "_build", "cusparse_baseline")


def as_oracle(M, dtype=np.float32):
    k, b, o, v = M.download()
    return O.OracleMatrix(M.num_rows, M.num_cols, k, b, o, v.astype(dtype), M.transposed)


def write_csr(path, nr, nc, rp, ci, v):
    with open(path, "wb") as f:
        np.array([nr, nc, ci.size], np.int64).tofile(f); rp.astype(np.int32).tofile(f); ci.astype(np.int32).tofile(f); v.astype(np.float32).tofile(f)


def cusparse(kind, csr, reps=3):
    if not os.path.exists(CUSP):
        return None
    out = subprocess.run([CUSP, kind, csr, str(reps)], capture_output=True, text=True, timeout=3600)
    for l in out.stdout.splitlines():
        if l.startswith("CUSPARSE_"):
            t = l.split(); return {"ms": float(t[1]), "raw": l}
    return {"error": (out.stdout + out.stderr)[-400:]}


def ours_spmv(A, x, reps=50):
    y = torch.empty(A.num_rows, device="cuda")
    for _ in range(5): B.bmSparse_SpMV(A, x, y)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): B.bmSparse_SpMV(A, x, y)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, y


def ours_spgemm(A, Bt, reps=3):
    best = 1e30; info = None
    for i in range(reps + 1):
        torch.cuda.synchronize(); t = time.perf_counter()
        C, info = B.bmSparse_mult(A, Bt, None, 0, True, 5)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
        if i: best = min(best, dt)
        cb, cn = C.block_num, C.nnz
        del C
    return best, info, cb, cn


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--tag", default="r1"); ap.add_argument("--only", default="")
    a = ap.parse_args()
    res = {}
    tmp = tempfile.mkdtemp(prefix="bmsp_cmp_")
    d = lambda t: torch.from_numpy(t).cuda()
    want = lambda n: (not a.only) or n in a.only.split(",")

    if want("spmv_p4096"):
        nr, nc, rp, ci, v = G.poisson5pt(4096, 4096)
        A = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v))
        ms, y = ours_spmv(A, torch.ones(nc, device="cuda"))
        nbytes = A.spmv_bytes()
        r = {"rows": nr, "nnz": int(ci.size), "blocks": A.block_num, "algorithmic_bytes": nbytes,
             "ours_fp16": {"ms": ms, "GBps": nbytes / ms / 1e6}}
        A32 = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v), dtype=torch.float32)
        ms32, y32 = ours_spmv(A32, torch.ones(nc, device="cuda"))
        r["ours_fp32_matrix"] = {"ms": ms32}
        if O.ref_cuda_bin("ref_spmv"):
            yr, us = O.run_ref_spmv(as_oracle(A32), tmp, reps=5)
            r["reference_bmsparse_cuda_fp32"] = {"ms": us / 1e3, "note": "bmSparse_SpMV<float,float> incl. its per-call reduce_by_key+scan and cudaDeviceSynchronize (SPMV.cu:191-230), best of 5",
                                                 "matches_ours": bool(np.array_equal(yr, y32.cpu().numpy()))}
        csr = os.path.join(tmp, "p4096.csr"); write_csr(csr, nr, nc, rp, ci, v)
        r["cusparse_csr_fp32"] = cusparse("spmv", csr, 20)
        res["spmv_p4096"] = r; print(json.dumps({"spmv_p4096": r}), flush=True)
        del A, A32

    for name, gen in (("spgemm_p256", lambda: G.poisson5pt(256, 256)), ("spgemm_u1m", lambda: G.uniform_random(1_000_000, 16, seed=2)),
                      ("spgemm_bc4m", lambda: G.block_clustered(524288)), ("spgemm_p4096", lambda: G.poisson5pt(4096, 4096)),
                      ("spgemm_rmat16", lambda: G.rmat(16))):
        if not want(name):
            continue
        nr, nc, rp, ci, v = gen()
        ref_valid = name != "spgemm_rmat16"      # R-MAT has empty 8-row block rows: the reference indexes a compacted array and breaks
        A = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v)); Bt = B.bmSpMatrix.from_csr(nr, nc, d(rp), d(ci), d(v), transpose=True)
        flops = 2 * int(np.diff(rp).astype(np.int64)[ci].sum())
        ms, info, cb, cn = ours_spgemm(A, Bt)
        r = {"rows": nr, "nnz": int(ci.size), "flops": flops, "c_blocks": cb, "c_nnz": cn,
             "ours": {"ms": ms, "GFLOPs": flops / ms / 1e6, "symbolic_ms": info.symbolic_ms, "numeric_ms": info.numeric_ms, "numeric_path": info.numeric_path}}
        if not ref_valid:
            r["reference_bmsparse_cuda"] = {"skipped": "input has empty block rows: the reference cannot produce a valid answer (SURVEY Appendix B)"}
        elif O.ref_cuda_bin("ref_spgemm"):
            try:
                oa, ob = as_oracle(A, np.float16), as_oracle(Bt, np.float16)
                for tc in (5, 4):
                    ref, us = O.run_ref_spgemm(oa, ob, tmp, tc_version=tc, mode=0, reps=2, keep=True)
                    r[f"reference_bmsparse_cuda_tc{tc}"] = {"ms": us / 1e3, "GFLOPs": flops / us / 1e3, "c_blocks": int(ref.block_num), "c_nnz": int(ref.nnz),
                                                            "note": "bmSparse_mult<half,float> whole call as its main times it (SPGEMM.cu:1274-1280), best of 2"}
                    del ref
            except Exception as e:
                r["reference_bmsparse_cuda"] = {"error": str(e)[-300:]}
        csr = os.path.join(tmp, name + ".csr"); write_csr(csr, nr, nc, rp, ci, v)
        cs = cusparse("spgemm", csr, 2)
        if cs and "ms" in cs: cs["GFLOPs"] = flops / cs["ms"] / 1e6
        r["cusparse_spgemm_fp32"] = cs
        res[name] = r; print(json.dumps({name: r}), flush=True)
        del A, Bt
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"compare_{a.tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
