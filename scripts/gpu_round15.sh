#!/bin/bash
# last call of round 1 (3.6 GPU-minutes left): the GPU suite including the new parity test against the reference's own
# CUDA backend, that backend timed as a second baseline, and a reduced-size bench.py pass (e2e path included)
set -u
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -q -rs -s > gpurun_out/pytest_gpu_r15.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_r15.log
grep -h "product vs reference CUDA" gpurun_out/pytest_gpu_r15.log
timeout 45 python scripts/ref_gpu_baseline.py --reps 10 > gpurun_out/ref_gpu_baseline.json 2> gpurun_out/ref_gpu_baseline.err; echo "ref_gpu_baseline rc=$?"
tail -c 1500 gpurun_out/ref_gpu_baseline.json; tail -3 gpurun_out/ref_gpu_baseline.err
timeout 60 python bench.py --o-segs 20,20 --v-segs 50,50,50,50 --steps 1 --warmup 1 --cpu-dests 1 > gpurun_out/bench_small_r15.json 2> gpurun_out/bench_small_r15.err; echo "bench small rc=$?"
tail -c 600 gpurun_out/bench_small_r15.json; tail -2 gpurun_out/bench_small_r15.err
