#!/bin/bash
# final measurements of a round: `bash tools/gpu_final.sh N` on a box with N GPUs -> gpurun_out/final_bench_nN.json (+ tests at N <= 4)
cd "$(dirname "$0")/.."
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest_n1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_n1.log
  tail -4 gpurun_out/final_pytest_n1.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
  ( time python bench.py --steps 200 --warmup 20 ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_n1_ref.json 2>> gpurun_out/final_bench_n1.err
  # profiler passes (numbers printed under ncu are never bench values): launch list of the bench command, full capture of the SpMV kernel
  # and of the conversion kernels
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 20 --warmup 5 --no-strong --no-cpu --no-spgemm > gpurun_out/final_ncu_bench.log 2>&1
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel -c 1 -f -o gpurun_out/final_stream python tools/spmv_bench.py p4096 3 > gpurun_out/final_ncu_stream.log 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:"chunk_row_kernel|rank_kernel|head_count_kernel|emit_kernel|values_kernel|derive_|check_rowptr" -c 8 -f -o gpurun_out/final_convert python tools/spmv_bench.py p4096 1 > gpurun_out/final_ncu_convert.log 2>&1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  if [ "$N" -le 4 ]; then timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/final_pytest_n$N.log 2>&1; tail -3 gpurun_out/final_pytest_n$N.log; fi
  ( time $TR --master-port 29600 bench.py --gpus $N --steps 200 --warmup 20 ) > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc=$?"
  tail -3 gpurun_out/final_bench_n$N.err
fi
python - <<PY
import json
d=json.load(open('gpurun_out/final_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('skipped'))
for k,v in d.get('spgemm',{}).get('configs',{}).items(): print(k, {q:v.get(q) for q in ('ms','gflops','reference_cuda_ms','cusparse_ms','faster_than_both','error')}, v.get('roofline',{}).get('frac'))
print(d.get('convert',{}).get('ms'), d.get('convert',{}).get('roofline',{}).get('frac'))
s=d.get('strong',{})
for k,v in s.items():
    if isinstance(v,dict): print(k, v.get('value'), v.get('ms_per_step', v.get('ms')), v.get('halo'))
PY
