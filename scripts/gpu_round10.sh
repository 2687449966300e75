#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 1 -c 1 -f -o gpurun_out/prof_contract_skinny \
  python scripts/ncu_gemm.py skinny 256 > gpurun_out/ncu_skinny.log 2>&1; echo "ncu skinny rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 1 -c 1 -f -o gpurun_out/prof_contract_outer \
  python scripts/ncu_gemm.py outer 256 > gpurun_out/ncu_outer.log 2>&1; echo "ncu outer rc=$?"
