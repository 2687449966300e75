#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lowint.py -x -q 2>&1 | tail -3
for p in "ab=acde*bcde vvovo 100" "ab=cade*bced vvoov 100" "ab=acde*cbed vvovo 100"; do
  set -- $p
  SIPGPU_LOWINT_SLAB=1 timeout 120 python scripts/ncu_pattern.py "$1" $2 $3
done
bash scripts/gpu_ncu_pattern.sh "ab=acde*bcde" vvovo 100 slab77_vvovo 2>&1 | tail -1
