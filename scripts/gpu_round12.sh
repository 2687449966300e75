#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workload.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== D-friendly order OFF"; SIPGPU_DFAST_ORDER=0 timeout 800 python scripts/sweep_patterns.py 2>&1 | grep "^ranks"; cp gpurun_out/sweep_patterns.json gpurun_out/sweep_patterns_dfast0.json
echo "== D-friendly order ON"; SIPGPU_DFAST_ORDER=1 timeout 800 python scripts/sweep_patterns.py 2>&1 | tail -22
