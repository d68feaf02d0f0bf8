#!/bin/bash
# the whole GPU suite and smoke() on one GPU: bash tools/gpu_pytest_only.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest_n1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_n1.log
tail -5 gpurun_out/final_pytest_n1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
