#!/bin/bash
# One gpurun call: GPU parity tests, smoke, ncu launch list of a reduced CCSD step, ncu --set full captures of the
# contraction (ring chain, 8192^3 GEMM, small-block batch) and permute kernels, then the default N=1 bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_ccsd_small.csv \
  python bench.py --o-segs 20,20 --v-segs 50,50,50,50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 1 -c 2 -f -o gpurun_out/prof_contract_ring \
  python scripts/ncu_gemm.py ring 296 > gpurun_out/ncu_ring.log 2>&1; echo "ncu ring rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 1 -c 1 -f -o gpurun_out/prof_contract_gemm \
  python scripts/ncu_gemm.py gemm 8192 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 1 -c 1 -f -o gpurun_out/prof_contract_small \
  python scripts/ncu_gemm.py small 16 > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute -c 6 -f -o gpurun_out/prof_permute \
  python scripts/ncu_gemm.py permute 64 > gpurun_out/ncu_permute.log 2>&1; echo "ncu permute rc=$?"
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -c 1200 gpurun_out/bench_ref.json
tail -5 gpurun_out/pytest_gpu.log
