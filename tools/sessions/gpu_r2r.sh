#!/bin/bash
# round-2 session r (1 GPU): producer work ahead of the stage wait; ncu captures for profiles/ (U1M passes, BC4M dense pass, launch list)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/spmv_bench.py p4096 300 2>&1 | tail -1 | tee gpurun_out/r2r_spmv.log
python tools/spmv_bench.py p2048 300 2>&1 | tail -1 | tee -a gpurun_out/r2r_spmv.log
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2r_halo.log
timeout 300 python -m pytest tests/test_gpu_spmv.py -x -q 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spgemm_pass_kernel --launch-skip 3 -c 3 -f -o gpurun_out/r2r_u1m python tools/spgemm_bench.py u1m --reps 1 > gpurun_out/r2r_ncu_u1m.log 2>&1
tail -1 gpurun_out/r2r_ncu_u1m.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"spgemm_pass_kernel|spgemm_dense_kernel" --launch-skip 3 -c 3 -f -o gpurun_out/r2r_bc4m python tools/spgemm_bench.py bc4m --reps 1 > gpurun_out/r2r_ncu_bc4m.log 2>&1
tail -1 gpurun_out/r2r_ncu_bc4m.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel -c 1 -f -o gpurun_out/r2r_stream python tools/spmv_bench.py p4096 3 > gpurun_out/r2r_ncu_stream.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 20 --warmup 5 --no-strong --no-cpu --no-spgemm > gpurun_out/r2r_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2r_ncu_bench.log
