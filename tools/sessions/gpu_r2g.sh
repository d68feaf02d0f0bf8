#!/bin/bash
# round-2 session g (2 GPUs): R-MAT-22 SpMV strong scaling with coalesced peer stores, halo step overhead
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python tools/rmat_scale.py --what spmv,poisson --steps 100 > gpurun_out/r2g_rmat_n1.json 2> gpurun_out/r2g.err
$TR --master-port 29521 tools/rmat_scale.py --what spmv,poisson --steps 100 > gpurun_out/r2g_rmat_n2.json 2>> gpurun_out/r2g.err
cat gpurun_out/r2g_rmat_n1.json gpurun_out/r2g_rmat_n2.json | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
$TR --master-port 29522 bench.py --gpus 2 --steps 100 --warmup 10 --no-strong > gpurun_out/r2g_bench2.json 2>> gpurun_out/r2g.err
python -c "
import json
d=json.load(open('gpurun_out/r2g_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e'])
"
tail -3 gpurun_out/r2g.err
