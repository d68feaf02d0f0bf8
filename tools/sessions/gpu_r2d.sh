#!/bin/bash
# round-2 session d: streaming SpMV kernel with one producer warp per group -- geometry sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_cpp_shim.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
run() { echo "== $*" >> gpurun_out/r2d_spmv.log; env "$@" timeout 120 python tools/spmv_bench.py ${W:-p4096} 300 2>&1 | tail -1 >> gpurun_out/r2d_spmv.log; }
run BMSP_SPMV_KERNEL=1
for c in 17 16 15 14 23 22 32 61; do run BMSP_SPMV_CFG=$c; done
run BMSP_SPMV_CFG=16 BMSP_SPMV_STAGES=1
run BMSP_SPMV_CFG=14 BMSP_SPMV_STAGES=2
run BMSP_SPMV_CFG=22 BMSP_SPMV_STAGES=2
W=bc run BMSP_SPMV_KERNEL=1
W=bc run BMSP_SPMV_CFG=0
W=p2048 run BMSP_SPMV_CFG=0
cat gpurun_out/r2d_spmv.log
