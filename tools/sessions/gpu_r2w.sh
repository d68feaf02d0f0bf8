#!/bin/bash
# round-2 session w (N GPUs): measured-cost split of the R-MAT-22 SpMV, warp-per-row fix-up; usage: bash tools/gpu_r2w.sh N
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" -le 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmv.py -x -q > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
  python tools/rmat_scale.py --what spmv --steps 100 > gpurun_out/r2w_rmat_n1.json 2> gpurun_out/r2w.err; cut -c1-400 gpurun_out/r2w_rmat_n1.json
fi
$TR --master-port 29561 tools/rmat_scale.py --what spmv --steps 200 > gpurun_out/r2w_rmat_n$N.json 2>> gpurun_out/r2w.err
cut -c1-1800 gpurun_out/r2w_rmat_n$N.json
tail -3 gpurun_out/r2w.err
