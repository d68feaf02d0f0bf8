#!/bin/bash
# round-2 session e: full GPU test suite (incl. full-size parity), SpMV sweep, ncu capture, the new bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -25 gpurun_out/r2e_pytest.log
run() { echo "== $*" >> gpurun_out/r2e_spmv.log; env "$@" timeout 120 python tools/spmv_bench.py ${W:-p4096} 300 2>&1 | tail -1 >> gpurun_out/r2e_spmv.log; }
run BMSP_SPMV_CTAS=0
run BMSP_SPMV_CTAS=6
run BMSP_SPMV_CTAS=5 BMSP_SPMV_STAGES=2
W=bc run BMSP_SPMV_KERNEL=1
W=bc run BMSP_SPMV_CTAS=0
W=bc run BMSP_SPMV_RT=32
W=bc run BMSP_SPMV_RT=64
W=p2048 run BMSP_SPMV_CTAS=0
cat gpurun_out/r2e_spmv.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel -c 1 -f -o gpurun_out/r2e_stream python tools/spmv_bench.py p4096 3 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2e_bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2e_bench_ref.json 2>> gpurun_out/r2e_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','verified')}, d['roofline']['frac'], d['e2e']['value'], d.get('skipped'))
for k,v in d.get('spgemm',{}).get('configs',{}).items(): print(k, {q:v.get(q) for q in ('ms','gflops','reference_cuda_ms','cusparse_ms','faster_than_both','error')}, v.get('roofline',{}).get('frac'))
print(d.get('convert')); print(json.dumps(d.get('strong'))[:1500])
"
