#!/bin/bash
# round-2 session q (2 GPUs): tile index through the stage header; bundles at N = 2; weak scaling
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python tools/spmv_bench.py p4096 300 2>&1 | tail -1 | tee gpurun_out/r2q_spmv.log
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2q_halo.log
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmv.py -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
python tools/rmat_scale.py --what spmv --steps 100 > gpurun_out/r2q_rmat_n1.json 2> gpurun_out/r2q.err
$TR --master-port 29551 tools/rmat_scale.py --what spmv,poisson --steps 100 > gpurun_out/r2q_rmat_n2.json 2>> gpurun_out/r2q.err
cat gpurun_out/r2q_rmat_n1.json gpurun_out/r2q_rmat_n2.json | cut -c1-700
$TR --master-port 29552 bench.py --gpus 2 --steps 100 --warmup 10 --no-strong > gpurun_out/r2q_bench2.json 2>> gpurun_out/r2q.err
python -c "
import json
d=json.load(open('gpurun_out/r2q_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'])
"
tail -2 gpurun_out/r2q.err
