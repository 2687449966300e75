#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/tune_permute.py 2>&1 | tail -1 | tee gpurun_out/tune_permute.json
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 2200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
