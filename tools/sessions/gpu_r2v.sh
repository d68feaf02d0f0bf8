#!/bin/bash
# round-2 session v (1 GPU): the shards of an 8-way R-MAT-22 SpMV, timed alone; kernel durations of the shard products
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/shard_probe.py 8 100 2>&1 | tail -12 | tee gpurun_out/r2v_shards.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches.csv -k regex:"spmv_|split|fix|halo" python tools/shard_probe.py 8 2 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
