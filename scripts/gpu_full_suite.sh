#!/bin/bash
# the whole GPU suite + smoke (+ the pattern sweep when SWEEP=1): the end-of-round check
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -24 > gpurun_out/pytest_gpu_full.txt
cat gpurun_out/pytest_gpu_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
if [ -n "$SWEEP" ]; then
  SWEEP_OUT=sweep_patterns_final.json timeout 900 python scripts/sweep_patterns.py > gpurun_out/sweep_patterns_final.txt 2>&1
  tail -22 gpurun_out/sweep_patterns_final.txt
fi
