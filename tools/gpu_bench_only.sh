#!/bin/bash
# the bench line and the reference arm only (N = 1): bash tools/gpu_bench_only.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python bench.py --steps 200 --warmup 20 ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_n1_ref.json 2>> gpurun_out/final_bench_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/final_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d.get('skipped'))
for k,v in d.get('spgemm',{}).get('configs',{}).items(): print(k, {q:v.get(q) for q in ('ms','reference_cuda_ms','cusparse_ms','faster_than_both','error')})
PY
