#!/bin/bash
# re-capture the ncu evidence with the final kernels of the round; the reports are summarised ON the GPU box
# (scripts/summarize_ncu.py) and deleted, because gpurun_out/ only travels back when it is < 64 MiB
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_ccsd_small.csv \
  python bench.py --o-segs 20,20 --v-segs 50,50,50,50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
python scripts/summarize_ncu.py list gpurun_out/launches_ccsd_small.csv > gpurun_out/r01_launches_ccsd_small.txt; rm -f gpurun_out/launches_ccsd_small.csv
for cfg in "ring 296 contract_kernel 1 contract_ring" "gemm 8192 contract_kernel 1 contract_gemm" "skinny 256 contract_kernel 1 contract_skinny" "permute 64 permute 6 permute"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c $4 -f -o gpurun_out/prof_$5 \
    python scripts/ncu_gemm.py $1 $2 > gpurun_out/ncu_$5.log 2>&1; echo "ncu $5 rc=$?"
  python scripts/summarize_ncu.py rep gpurun_out/prof_$5.ncu-rep > gpurun_out/r01_ncu_$5.txt 2>/dev/null
  rm -f gpurun_out/prof_$5.ncu-rep
done
ls -la gpurun_out
