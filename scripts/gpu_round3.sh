#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/tune_permute.py 2>&1 | tail -1 | tee gpurun_out/tune_permute.json
timeout 1500 python scripts/sweep.py "$@" > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -25 gpurun_out/sweep.log
