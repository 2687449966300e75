#!/bin/bash
# work-list / SIAL front-end GPU tests, the op-at-a-time vs recorded comparison, and the DRAM traffic of the contraction
# launches of a full-size CCSD step (single-pass ncu metrics, no replay).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_worklist.py -x -q > gpurun_out/pytest_wl.log 2>&1; echo "pytest wl rc=$?"; tail -15 gpurun_out/pytest_wl.log
timeout 600 python scripts/sial_frontend_bench.py 20,20 50,50,50,50 2 > gpurun_out/sial_frontend_bench.json 2> gpurun_out/sial_frontend_bench.err; echo "frontend bench rc=$?"
tail -c 1500 gpurun_out/sial_frontend_bench.json; tail -3 gpurun_out/sial_frontend_bench.err
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:contract_kernel --csv \
  --log-file gpurun_out/traffic_ccsd_full.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_traffic.log 2>&1
echo "ncu traffic rc=$?"; tail -c 600 gpurun_out/traffic_ccsd_full.csv
