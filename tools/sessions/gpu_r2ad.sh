#!/bin/bash
# round-2 session ad (1 GPU, short): tile-claim counters allocated by the plan
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 40 python tools/spmv_bench.py p4096 100 2>&1 | tail -1 | tee gpurun_out/r2ad.log
