#!/bin/bash
# round-2 session b: streaming SpMV kernel -- tests, A/B sweep, sanitizer, ncu
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_convert.py tests/test_gpu_cpp_shim.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
run() { echo "== $*" >> gpurun_out/r2b_spmv.log; env "$@" timeout 120 python tools/spmv_bench.py ${W:-p4096} 300 2>&1 | tail -1 >> gpurun_out/r2b_spmv.log; }
run BMSP_SPMV_KERNEL=1
run BMSP_SPMV_NG=4
run BMSP_SPMV_NG=4 BMSP_SPMV_STAGES=4
run BMSP_SPMV_NG=4 BMSP_SPMV_STAGES=5
run BMSP_SPMV_NG=4 BMSP_SPMV_STAGES=6
run BMSP_SPMV_NG=7
run BMSP_SPMV_NG=3
run BMSP_SPMV_NG=2
run BMSP_SPMV_NG=4 BMSP_SPMV_RT=32
W=bc run BMSP_SPMV_KERNEL=1
W=bc run BMSP_SPMV_NG=4
cat gpurun_out/r2b_spmv.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -k "random or poisson_and or overflow" -x -q > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2b_memcheck.log
tail -4 gpurun_out/r2b_memcheck.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel -c 1 -f -o gpurun_out/r2b_stream python tools/spmv_bench.py p4096 3 > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_ncu.log
