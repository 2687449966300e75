#!/bin/bash
# compute-sanitizer over a slice of the GPU parity tests: memcheck (out-of-bounds / misaligned accesses of the
# predicated cp.async gathers, narrow tiles, split-K) and racecheck (shared-memory hazards of the mbarrier ring)
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "ragged or split_k or sial_cc_patterns or permute_batched or contract_sliced or chained or small_test" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "small_test2 or transpose4d or split_k" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/sanitize_racecheck.log
