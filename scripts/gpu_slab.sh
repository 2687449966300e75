#!/bin/bash
mkdir -p gpurun_out
SIPGPU_SLAB_MINK=16 timeout 900 python -m pytest tests/test_gpu_lowint.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
SWEEP_RANKS=222 SIPGPU_SLAB_MINK=1024 SWEEP_OUT=sweep_mink1024.json timeout 600 python scripts/sweep_patterns.py 2>&1 | tail -14
SWEEP_RANKS=222 SIPGPU_SLAB_MINK=16 SWEEP_OUT=sweep_mink16.json timeout 600 python scripts/sweep_patterns.py 2>&1 | tail -14
