#!/bin/bash
# round-2 session y (1 GPU): local_row by halving steps -- SpGEMM tests and the four configs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2 | tee gpurun_out/r2y.log
for m in p4096 u1m bc4m p256; do python tools/spgemm_bench.py $m --reps 3 2>&1 | tail -1 | tee -a gpurun_out/r2y.log; done
python tools/spgemm_bench.py rmat18 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/r2y.log
