#!/bin/bash
# round-2 session aa (1 GPU): pair list sorted by class in FILL, light pairs one lane each in NUMERIC
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2aa.log
timeout 400 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_fullsize.py tests/test_gpu_reference_cuda.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2aa.log
for s in 1 0; do
  for m in p4096 p1024; do BMSP_SPGEMM_SORT=$s python tools/spgemm_bench.py $m --reps 3 2>&1 | tail -1 | tee -a gpurun_out/r2aa.log; done
done
