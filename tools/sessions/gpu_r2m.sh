#!/bin/bash
# round-2 session m (2 GPUs): split halo launch (boundary / interior), block-parallel halo by push kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2m_halo.log
BMSP_HALO_SPLIT=0 python tools/halo_selftest.py 300 2>&1 | tail -1 | tee -a gpurun_out/r2m_halo.log
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
python tools/rmat_scale.py --what spmv,poisson --steps 100 > gpurun_out/r2m_rmat_n1.json 2> gpurun_out/r2m.err
$TR --master-port 29531 tools/rmat_scale.py --what spmv,poisson --steps 100 > gpurun_out/r2m_rmat_n2.json 2>> gpurun_out/r2m.err
cat gpurun_out/r2m_rmat_n1.json gpurun_out/r2m_rmat_n2.json | cut -c1-700
$TR --master-port 29532 bench.py --gpus 2 --steps 100 --warmup 10 --no-strong > gpurun_out/r2m_bench2.json 2>> gpurun_out/r2m.err
python -c "
import json
d=json.load(open('gpurun_out/r2m_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'])
"
tail -2 gpurun_out/r2m.err
