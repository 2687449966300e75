#!/bin/bash
mkdir -p gpurun_out
SWEEP_OUT=sweep_patterns_slab.json timeout 900 python scripts/sweep_patterns.py > gpurun_out/sweep_patterns_slab.txt 2>&1
tail -22 gpurun_out/sweep_patterns_slab.txt
timeout 900 python -m pytest tests/test_gpu_lowint.py tests/test_gpu_parity.py tests/test_gpu_worklist.py tests/test_gpu_z_cross_product.py -x -q 2>&1 | tail -3
