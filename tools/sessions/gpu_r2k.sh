#!/bin/bash
# round-2 session k (1 GPU): fine index with prefetch + deferred long buckets, dense-block numeric kernel, halo flag through the header
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_reference_cuda.py tests/test_gpu_fullsize.py tests/test_gpu_spmv.py -x -q ) > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -6 gpurun_out/r2k_pytest.log
for w in u1m rmat18 bc4m p4096; do python tools/spgemm_bench.py $w --reps 3 2>&1 | tail -1 >> gpurun_out/r2k_spgemm.log; done
BMSP_SPGEMM_DENSE=0 python tools/spgemm_bench.py bc4m --reps 2 2>&1 | tail -1 >> gpurun_out/r2k_spgemm.log
BMSP_SPGEMM_FINE=0 python tools/spgemm_bench.py rmat18 --reps 2 2>&1 | tail -1 >> gpurun_out/r2k_spgemm.log
cat gpurun_out/r2k_spgemm.log
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2k_halo.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spgemm_dense_kernel -c 1 -f -o gpurun_out/r2k_bc4m python tools/spgemm_bench.py bc4m --reps 0 > gpurun_out/r2k_ncu_bc4m.log 2>&1
tail -2 gpurun_out/r2k_ncu_bc4m.log
