#!/bin/bash
# round-2 session u (1 GPU): share of claimed tiles 12 / 18 / 25 / 33 %; conversion with the chunk-row table
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2u.log
for d in 12 18 25 33; do
  echo "DYN=$d" | tee -a gpurun_out/r2u.log
  BMSP_SPMV_DYN=$d timeout 120 python tools/spmv_bench.py p4096 300 2>&1 | tail -1 | tee -a gpurun_out/r2u.log
  BMSP_SPMV_DYN=$d timeout 120 python tools/spmv_bench.py p2048 300 2>&1 | tail -1 | tee -a gpurun_out/r2u.log
  BMSP_SPMV_DYN=$d HALO_ONLY=1 timeout 120 python tools/halo_selftest.py 300 2>&1 | tail -1 | tee -a gpurun_out/r2u.log
done
timeout 600 python -m pytest tests/test_gpu_convert.py tests/test_gpu_spgemm.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2u.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-strong --no-cpu --no-spgemm > gpurun_out/r2u_bench.json 2> gpurun_out/r2u.err
python -c "
import json
d=json.load(open('gpurun_out/r2u_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('convert'))
" | tee -a gpurun_out/r2u.log
