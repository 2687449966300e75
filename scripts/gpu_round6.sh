#!/bin/bash
# work-list GPU tests (after the chain fix) and the C++ op-at-a-time vs recorded pardo comparison at three block sizes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_worklist.py -x -q > gpurun_out/pytest_wl.log 2>&1; echo "pytest wl rc=$?"; tail -15 gpurun_out/pytest_wl.log
( cd scripts/micro && g++ -O2 -std=c++17 wl_pardo_bench.cpp -I../../include -L../../aces4_b200/lib -lsipgpu -Wl,-rpath,'$ORIGIN/../../aces4_b200/lib' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -o wl_pardo_bench ) || echo "build failed"
: > gpurun_out/wl_pardo_bench.jsonl
for cfg in "16 3 16 6" "8 4 24 6" "20 2 50 4" "20 3 50 6"; do
  timeout 600 scripts/micro/wl_pardo_bench $cfg 3 >> gpurun_out/wl_pardo_bench.jsonl 2>> gpurun_out/wl_pardo_bench.err; echo "cfg $cfg rc=$?"
done
cat gpurun_out/wl_pardo_bench.jsonl; tail -3 gpurun_out/wl_pardo_bench.err
