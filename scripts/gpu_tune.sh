#!/bin/bash
# kernel tuning call: GPU parity tests on the default build, then the tuning probe for every configuration given
# (default | w16 | <alternative lib under aces4_b200/lib>)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "$@"; do
  unset SIPGPU_LIB SIPGPU_WARPS16
  case "$cfg" in
    default) ;;
    w16) export SIPGPU_WARPS16=1 ;;
    *) export SIPGPU_LIB=$PWD/aces4_b200/lib/$cfg ;;
  esac
  echo "== $cfg" | tee -a gpurun_out/tune.log
  timeout 600 python scripts/tune_contract.py 2>&1 | tail -2 | tee -a gpurun_out/tune.log
done
