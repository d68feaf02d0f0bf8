"""Markdown tables for DESIGN.md / README.md from the committed bench lines (profiles/r2_bench_n{1,2,4,8}.json):
python tools/results_tables.py [spgemm|multi|readme]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(n):
    p = os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")
    return json.load(open(p)) if os.path.exists(p) else None


TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))


def spgemm():
    d = load(1)
    names = {"p256": "P256 (config 1 shape)", "p4096": "P4096", "u1m": "U1M (config 3)", "bc4m": "BC4M (config 4)"}
    out = ["| config | ours ms (symbolic + numeric) | GFLOP/s | HBM roofline frac (algorithmic bytes) | DRAM traffic (ncu) | reference bmSparse CUDA ms | cuSPARSE ms | vs reference | vs cuSPARSE |",
           "|---|---|---|---|---|---|---|---|---|"]
    for k in ("p256", "p4096", "u1m", "bc4m"):
        v = d["spgemm"]["configs"][k]
        r = v["roofline"]
        t = r.get("traffic") or TRAFFIC.get(f"spgemm_{k}_dram_bytes")        # captures made after the bench line was written
        tr = (f"{t / 1e9:.1f} GB" if t >= 1e8 else f"{t / 1e6:.1f} MB") if t else "—"
        alg = v.get("cusparse", {}).get("alg")
        out.append(f"| {names[k]} | **{v['ms']:.2f}** ({v['symbolic_ms']:.2f} + {v['numeric_ms']:.2f}, {v['numeric_path']}) | {v['gflops']:.1f} | "
                   f"{r['frac']:.3f} ({r['algorithmic_bytes'] / 1e9:.2f} GB) | {tr} | {v['reference_cuda_ms']:.1f} | {v['cusparse_ms']:.1f} (ALG{alg}) | "
                   f"{v['reference_cuda_ms'] / v['ms']:.1f}× | {v['cusparse_ms'] / v['ms']:.1f}× |")
    return "\n".join(out)


def multi():
    ds = {n: load(n) for n in (1, 2, 4, 8)}
    ds = {n: d for n, d in ds.items() if d}
    base = ds[1]
    out = ["| N GPUs | weak: P4096 slab per GPU, µs / step | aggregate GB/s | efficiency | verified | e2e GB/s (ms / step) | strong: P4096 split N ways µs (speed-up) | RM22 SpMV µs (speed-up) | RM22 A·A s (speed-up) |",
           "|---|---|---|---|---|---|---|---|---|"]
    b = base["strong"]
    for n, d in sorted(ds.items()):
        s = d["strong"]
        ps, rs, rg = s["p4096_spmv_split"], s["rmat22_spmv"], s["rmat22_spgemm"]
        out.append(f"| {n} | {d['ms_per_step'] * 1e3:.1f} | {d['value']:.0f} | {d['value'] / (n * base['value']):.3f} | {d['verified']} | "
                   f"{d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.2f}) | {ps['ms_per_step'] * 1e3:.1f} ({b['p4096_spmv_split']['ms_per_step'] / ps['ms_per_step']:.2f}×) | "
                   f"{rs['ms_per_step'] * 1e3:.1f} ({b['rmat22_spmv']['ms_per_step'] / rs['ms_per_step']:.2f}×) | "
                   f"{rg['ms'] / 1e3:.2f} ({b['rmat22_spgemm']['ms'] / rg['ms']:.2f}×) |")
    return "\n".join(out)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "spgemm"
    print({"spgemm": spgemm, "multi": multi}[what]())
