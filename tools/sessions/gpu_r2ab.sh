#!/bin/bash
# round-2 session ab (1 GPU): ncu capture of the SpGEMM passes on P4096 and P256 with the class-sorted pair list (profiles/, traffic.json)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 150 ncu --set full --import-source on --clock-control none -k regex:"spgemm_pass_kernel" --launch-skip 3 -c 3 -f -o gpurun_out/r2ab_p4096 python tools/spgemm_bench.py p4096 --reps 1 > gpurun_out/r2ab_ncu.log 2>&1
timeout 60 ncu --set full --clock-control none -k regex:"spgemm_pass_kernel" --launch-skip 3 -c 3 -f -o gpurun_out/r2ab_p256 python tools/spgemm_bench.py p256 --reps 1 >> gpurun_out/r2ab_ncu.log 2>&1
tail -2 gpurun_out/r2ab_ncu.log
