#!/bin/bash
# round-2 session s (2 GPUs): halo path with one fence per pushing tile + acquire-load waits; bundle test; N = 2 bench with the strong section
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python tools/halo_selftest.py 300 2>&1 | tail -1 | tee gpurun_out/r2s_halo.log
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmv.py -x -q > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
( time $TR --master-port 29552 bench.py --gpus 2 --steps 200 --warmup 20 ) > gpurun_out/r2s_bench2.json 2> gpurun_out/r2s.err
python -c "
import json
d=json.load(open('gpurun_out/r2s_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'])
s=d.get('strong',{})
for k,v in s.items():
    if isinstance(v,dict): print(k, v.get('value'), v.get('ms_per_step', v.get('ms')), v.get('halo'))
"
tail -5 gpurun_out/r2s.err
