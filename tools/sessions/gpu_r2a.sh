#!/bin/bash
# round-2 session a: tests + A/B of the SpMV tile kernel (round-1 library vs tridiagonal fast path) + ncu capture
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
L=bmsparse_spgemm_spmv_b200/lib
for w in p4096 bc; do
  BMSP_LIB_PATH=$PWD/$L/libbmsparse_b200_r1.so python tools/spmv_bench.py $w 300 2>&1 | tail -1 | sed 's/^/r1  /' >> gpurun_out/r2a_spmv.log
  python tools/spmv_bench.py $w 300 2>&1 | tail -1 | sed 's/^/new /' >> gpurun_out/r2a_spmv.log
done
ncu --set full --import-source on --clock-control none -k regex:spmv_tile_kernel -c 1 -f -o gpurun_out/r2a_tile python tools/spmv_bench.py p4096 3 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_spmv.log
