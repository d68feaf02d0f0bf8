#!/bin/bash
# round-2 session z (1 GPU): lanes per surviving pair in the scalar NUMERIC pass (8 / 4 / 2) on the stencil configs
# (BMSP_SPGEMM_SPLIT, used below, was an experiment switch removed after this session: eight lanes per pair stayed)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2z.log
for s in 8 4 2; do
  echo "SPLIT=$s" | tee -a gpurun_out/r2z.log
  for m in p4096 p1024; do BMSP_SPGEMM_SPLIT=$s python tools/spgemm_bench.py $m --reps 3 2>&1 | tail -1 | tee -a gpurun_out/r2z.log; done
done
BMSP_SPGEMM_SPLIT=4 timeout 600 python -m pytest tests/test_gpu_spgemm.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2z.log
