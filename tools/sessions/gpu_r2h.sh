#!/bin/bash
# round-2 session h (1 GPU): new boundary / export tests, halo code-path self-test, final ncu capture of the streaming kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_convert.py tests/test_gpu_cpp_shim.py tests/test_gpu_spmv.py -x -q ) > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -15 gpurun_out/r2h_pytest.log
python tools/halo_selftest.py 300 2>&1 | tail -1 > gpurun_out/r2h_halo.log
BMSP_HALO_ROTATE=0 python tools/halo_selftest.py 300 2>&1 | tail -1 >> gpurun_out/r2h_halo.log
BMSP_HALO_FUSED=0 python tools/halo_selftest.py 300 2>&1 | tail -1 >> gpurun_out/r2h_halo.log
cat gpurun_out/r2h_halo.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel -c 1 -f -o gpurun_out/r2h_stream python tools/spmv_bench.py p4096 3 > gpurun_out/r2h_ncu.log 2>&1
tail -2 gpurun_out/r2h_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 5 --warmup 3 --no-strong --no-cpu --budget-s 60 > gpurun_out/r2h_ncu_bench.log 2>&1
tail -2 gpurun_out/r2h_ncu_bench.log | cut -c1-300
