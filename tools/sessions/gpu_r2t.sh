#!/bin/bash
# round-2 session t (1 GPU): tiles claimed at run time -- share of dynamically claimed tiles 0 / 25 / 50 / 100 %, plain and halo variant
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2t.log
for d in 0 25 50 100; do
  echo "DYN=$d" | tee -a gpurun_out/r2t.log
  BMSP_SPMV_DYN=$d timeout 120 python tools/spmv_bench.py p4096 300 2>&1 | tail -1 | tee -a gpurun_out/r2t.log
  BMSP_SPMV_DYN=$d HALO_ONLY=0 timeout 120 python tools/halo_selftest.py 300 2>&1 | tail -1 | tee -a gpurun_out/r2t.log
done
for d in 0 100; do
  BMSP_SPMV_DYN=$d timeout 120 python tools/spmv_bench.py p2048 300 2>&1 | tail -1 | tee -a gpurun_out/r2t.log
  BMSP_SPMV_DYN=$d BMSP_SPMV_KERNEL=2 timeout 120 python tools/spmv_bench.py bc 100 2>&1 | tail -1 | tee -a gpurun_out/r2t.log
done
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2t.log
