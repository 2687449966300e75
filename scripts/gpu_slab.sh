#!/bin/bash
mkdir -p gpurun_out
SIPGPU_SLAB_HYBRID=2 timeout 900 python -m pytest tests/test_gpu_lowint.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for p in "ab=acde*edcb ovooo 625" "ab=cdbe*aecd vovoo 250" "ab=acde*bedc vvovo 100"; do
  set -- $p
  SIPGPU_SLAB_HYBRID=1 timeout 120 python scripts/ncu_pattern.py "$1" $2 $3
  SIPGPU_SLAB_HYBRID=2 timeout 120 python scripts/ncu_pattern.py "$1" $2 $3
done
