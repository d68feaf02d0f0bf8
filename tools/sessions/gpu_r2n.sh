#!/bin/bash
# round-2 session n (1 GPU): CTA size of the fine-index passes, halo variant with its rare actions out of line, full test suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2n_halo.log
r() { echo "== $*" >> gpurun_out/r2n_spgemm.log; env "${@:2}" python tools/spgemm_bench.py $1 --reps 2 2>&1 | tail -1 >> gpurun_out/r2n_spgemm.log; }
r u1m BMSP_SPGEMM_T=256
r u1m BMSP_SPGEMM_T=512
r u1m BMSP_SPGEMM_T=1024
r rmat18 X=1
cat gpurun_out/r2n_spgemm.log | cut -c1-200
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -6 gpurun_out/r2n_pytest.log
