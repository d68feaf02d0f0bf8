#!/bin/bash
# round-2 session j (1 GPU): SpGEMM through B's fine index; halo variant of the streaming kernel with its multi-GPU code out of line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_reference_cuda.py tests/test_gpu_fullsize.py -x -q ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -6 gpurun_out/r2j_pytest.log
for w in u1m rmat18 p4096; do python tools/spgemm_bench.py $w --reps 3 2>&1 | tail -2 >> gpurun_out/r2j_spgemm.log; done
BMSP_SPGEMM_FINE=0 python tools/spgemm_bench.py u1m --reps 2 2>&1 | tail -1 >> gpurun_out/r2j_spgemm.log
cat gpurun_out/r2j_spgemm.log
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2j_halo.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spgemm_pass_kernel --launch-skip 3 -c 3 -f -o gpurun_out/r2j_u1m python tools/spgemm_bench.py u1m --reps 1 > gpurun_out/r2j_ncu_u1m.log 2>&1
tail -2 gpurun_out/r2j_ncu_u1m.log
