#!/bin/bash
# round-2 session c: streaming SpMV kernel after the producer / stage-ownership fixes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
run() { echo "== $*" >> gpurun_out/r2c_spmv.log; env "$@" timeout 120 python tools/spmv_bench.py ${W:-p4096} 300 2>&1 | tail -1 >> gpurun_out/r2c_spmv.log; }
run BMSP_SPMV_KERNEL=1
run BMSP_SPMV_NG=2
run BMSP_SPMV_NG=3
run BMSP_SPMV_NG=4
run BMSP_SPMV_NG=7
run BMSP_SPMV_NG=4 BMSP_SPMV_STAGES=4
run BMSP_SPMV_NG=2 BMSP_SPMV_STAGES=2
W=bc run BMSP_SPMV_KERNEL=1
W=bc run BMSP_SPMV_NG=4
W=p2048 run BMSP_SPMV_KERNEL=1
W=p2048 run BMSP_SPMV_NG=4
cat gpurun_out/r2c_spmv.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel -c 1 -f -o gpurun_out/r2c_stream python tools/spmv_bench.py p4096 3 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
