#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -30 > gpurun_out/pytest_gpu_full.txt
cat gpurun_out/pytest_gpu_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
