#!/bin/bash
set -x
mkdir -p gpurun_out
SWEEP_RANKS=242,244,255,233 SIPGPU_KORDER_BY_SIZE=0 SWEEP_OUT=sweep_korder0.json timeout 600 python scripts/sweep_patterns.py > gpurun_out/sweep_korder0.txt 2>&1
SWEEP_RANKS=242,244,255,233 SIPGPU_KORDER_BY_SIZE=1 SWEEP_OUT=sweep_korder1.json timeout 600 python scripts/sweep_patterns.py > gpurun_out/sweep_korder1.txt 2>&1
tail -18 gpurun_out/sweep_korder0.txt; tail -18 gpurun_out/sweep_korder1.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lowint.py tests/test_gpu_worklist.py -x -q 2>&1 | tail -4
