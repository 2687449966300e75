#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 800 python scripts/sweep_patterns.py 2>&1 | grep "^ranks"
timeout 1000 python scripts/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; grep "^contract\|^(T)\|^EOM" gpurun_out/sweep.log
timeout 300 python scripts/probe_gpu.py 2>&1 | tail -3
