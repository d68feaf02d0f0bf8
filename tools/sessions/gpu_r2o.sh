#!/bin/bash
# round-2 session o (2 GPUs): concurrent boundary launch, bundles of short block rows in the block-parallel SpMV
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2o_halo.log
BMSP_HALO_SPLIT=0 python tools/halo_selftest.py 300 2>&1 | tail -1 | tee -a gpurun_out/r2o_halo.log
for b in 0 1; do for w in u1m rmat22; do BMSP_SPMV_BUNDLES=$b python tools/spmv_bench.py $w 100 2>&1 | tail -1 | sed "s/^/bundles=$b /" | tee -a gpurun_out/r2o_spmv.log; done; done
timeout 300 python -m pytest tests/test_gpu_spmv.py -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
$TR --master-port 29541 tools/rmat_scale.py --what spmv,poisson --steps 100 > gpurun_out/r2o_rmat_n2.json 2> gpurun_out/r2o.err
cat gpurun_out/r2o_rmat_n2.json | cut -c1-700
$TR --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 10 --no-strong > gpurun_out/r2o_bench2.json 2>> gpurun_out/r2o.err
python -c "
import json
d=json.load(open('gpurun_out/r2o_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('numa'))
"
tail -2 gpurun_out/r2o.err
