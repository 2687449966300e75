#!/bin/bash
# memcheck over the kernels added late in round 2: slab_kernel (TMA-fed contractions), copy_bulk_kernel (block traffic)
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_lowint.py tests/test_gpu_worklist.py -x -q \
  -k "slab or section_of_gets or put_accumulate_stress" > gpurun_out/sanitize_memcheck_late.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck_late.log
