#!/bin/bash
# round-2 session p (1 GPU): bundles fix, where the halo entry point's extra microseconds come from
# (BMSP_HALO_FAKE, used below, was an experiment switch removed after this session)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmv.py -x -q 2>&1 | tail -2
for b in 0 1; do for w in rmat22 rmat20; do BMSP_SPMV_BUNDLES=$b python tools/spmv_bench.py $w 100 2>&1 | tail -1 | sed "s/^/bundles=$b /" | tee -a gpurun_out/r2p_spmv.log; done; done
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2p_halo.log
BMSP_HALO_FAKE=1 python tools/halo_selftest.py 300 2>&1 | tail -1 | tee -a gpurun_out/r2p_halo.log
BMSP_HALO_ROTATE=0 python tools/halo_selftest.py 300 2>&1 | tail -1 | tee -a gpurun_out/r2p_halo.log
HALO_ONLY=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_stream_kernel --launch-skip 20 -c 1 -f -o gpurun_out/r2p_halo python tools/halo_selftest.py 30 > gpurun_out/r2p_ncu_halo.log 2>&1
tail -1 gpurun_out/r2p_ncu_halo.log
