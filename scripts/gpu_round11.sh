#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workload.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 800 python scripts/sweep_patterns.py 2>&1 | tail -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 1 -c 1 -f -o gpurun_out/prof_contract_skinny \
  python scripts/ncu_gemm.py skinny 256 > gpurun_out/ncu_skinny.log 2>&1; echo "ncu skinny rc=$?"
