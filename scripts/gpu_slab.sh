#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lowint.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
bash scripts/gpu_ncu_pattern.sh "ab=cade*cbde" oovvo 250 slab_oovvo 2>&1 | tail -3
bash scripts/gpu_ncu_pattern.sh "ab=acde*bcde" vvovo 100 slab_vvovo 2>&1 | tail -3
