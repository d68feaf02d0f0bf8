#!/bin/bash
# round-2 session x (1 GPU): ncu capture of the three SpGEMM passes on P4096 (A*A of the 5-point stencil: 2.1 M tiny block rows)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/spgemm_bench.py p4096 --reps 3 2>&1 | tail -4 | tee gpurun_out/r2x_p4096.log
python tools/spgemm_bench.py p4096 --reps 3 --path 1 2>&1 | tail -2 | tee -a gpurun_out/r2x_p4096.log
for g in 8 32; do BMSP_SPGEMM_GROUP=$g python tools/spgemm_bench.py p4096 --reps 3 2>&1 | tail -1 | tee -a gpurun_out/r2x_p4096.log; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"spgemm_pass_kernel|rowinfo|finish_rows" --launch-skip 4 -c 5 -f -o gpurun_out/r2x_p4096 python tools/spgemm_bench.py p4096 --reps 1 > gpurun_out/r2x_ncu.log 2>&1
tail -2 gpurun_out/r2x_ncu.log
