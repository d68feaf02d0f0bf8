#!/bin/bash
# round-2 session ac (1 GPU, short): halo variant of the streaming kernel with the proxy fence behind the epoch wait (rank as its own peer)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 60 python tools/halo_selftest.py 200 2>&1 | tail -1 | tee gpurun_out/r2ac.log
