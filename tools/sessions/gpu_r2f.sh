#!/bin/bash
# round-2 session f (2 GPUs): multi-GPU parity tests + the bench line at N = 2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/spmv_bench.py p4096 300 2>&1 | tail -1 > gpurun_out/r2f_spmv.log
python tools/spmv_bench.py p2048 300 2>&1 | tail -1 >> gpurun_out/r2f_spmv.log
python tools/spmv_bench.py bc 100 2>&1 | tail -1 >> gpurun_out/r2f_spmv.log
cat gpurun_out/r2f_spmv.log
( time timeout 900 python -m pytest tests/test_gpu_dist.py -x -q ) > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -12 gpurun_out/r2f_pytest.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 ) > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2f_bench2.err
python -c "
import json
d=json.load(open('gpurun_out/r2f_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','verified','n_gpus')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d.get('skipped'))
print(json.dumps(d.get('strong'))[:2500])
"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2f_bench2_ref.json 2>> gpurun_out/r2f_bench2.err; cat gpurun_out/r2f_bench2_ref.json | cut -c1-600
