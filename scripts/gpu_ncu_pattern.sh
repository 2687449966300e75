#!/bin/bash
# $1 pattern $2 kinds $3 nb $4 tag [$5 lowint_scope]
mkdir -p gpurun_out
python scripts/ncu_pattern.py "$1" $2 $3 $5
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lowint|contract_kernel|slab_kernel" -s 1 -c 1 -f -o gpurun_out/ncu_$4 python scripts/ncu_pattern.py "$1" $2 $3 $5 > gpurun_out/ncu_$4.log 2>&1
tail -2 gpurun_out/ncu_$4.log
