#!/bin/bash
# round-2 session l (1 GPU): fine index second walk, dense kernel with column windows, halo tile intervals
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_reference_cuda.py tests/test_gpu_fullsize.py -x -q ) > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -6 gpurun_out/r2l_pytest.log
r() { echo "== $*" >> gpurun_out/r2l_spgemm.log; env "${@:2}" python tools/spgemm_bench.py $1 --reps 2 2>&1 | tail -1 >> gpurun_out/r2l_spgemm.log; }
r u1m X=1
r rmat18 X=1
r rmat18 BMSP_SPGEMM_FINE=0
r rmat20 X=1
r rmat20 BMSP_SPGEMM_FINE=0
r bc4m BMSP_SPGEMM_WCOLS=16
r bc4m BMSP_SPGEMM_WCOLS=8
r bc4m BMSP_SPGEMM_WCOLS=32
r bc4m BMSP_SPGEMM_DENSE=0
r p4096 X=1
cat gpurun_out/r2l_spgemm.log | cut -c1-200
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2l_halo.log
BMSP_SPGEMM_WCOLS=16 timeout 600 ncu --set full --import-source on --clock-control none -k regex:spgemm_dense_kernel -c 1 -f -o gpurun_out/r2l_bc4m python tools/spgemm_bench.py bc4m --reps 0 > gpurun_out/r2l_ncu_bc4m.log 2>&1
tail -1 gpurun_out/r2l_ncu_bc4m.log
