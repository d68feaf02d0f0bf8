#!/bin/bash
# round-2 session i (1 GPU): SpGEMM with 32-byte B records; ncu of the halo variant of the streaming kernel and of the U1M product
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_reference_cuda.py tests/test_gpu_fullsize.py -x -q ) > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
for w in u1m p4096 rmat18; do python tools/spgemm_bench.py $w --reps 3 2>&1 | tail -2 >> gpurun_out/r2i_spgemm.log; done
cat gpurun_out/r2i_spgemm.log
HALO_ONLY=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel --launch-skip 20 -c 1 -f -o gpurun_out/r2i_halo python tools/halo_selftest.py 30 > gpurun_out/r2i_ncu_halo.log 2>&1
tail -2 gpurun_out/r2i_ncu_halo.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spgemm_pass_kernel --launch-skip 3 -c 3 -f -o gpurun_out/r2i_u1m python tools/spgemm_bench.py u1m --reps 1 > gpurun_out/r2i_ncu_u1m.log 2>&1
tail -2 gpurun_out/r2i_ncu_u1m.log
