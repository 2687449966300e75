#!/bin/bash
set -u
mkdir -p gpurun_out
( cd scripts/micro && g++ -O2 -std=c++17 wl_pardo_bench.cpp -I../../include -L../../aces4_b200/lib -lsipgpu -Wl,-rpath,'$ORIGIN/../../aces4_b200/lib' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -o wl_pardo_bench ) || echo "build failed"
: > gpurun_out/wl_pardo_bench_idle.jsonl
for cfg in "16 3 16 6" "8 4 24 6" "32 2 32 4" "20 2 50 4" "20 3 50 6"; do
  timeout 600 scripts/micro/wl_pardo_bench $cfg 3 >> gpurun_out/wl_pardo_bench_idle.jsonl 2>> gpurun_out/wl_pardo_bench.err; echo "cfg $cfg rc=$?"
done
cat gpurun_out/wl_pardo_bench_idle.jsonl | cut -c150-900
timeout 600 python -m pytest tests/test_gpu_worklist.py -x -q 2>&1 | tail -3
