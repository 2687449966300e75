#!/bin/bash
# NVLink byte counters of the block-traffic kernels (both ranks under ncu): $1 = kernel regex, $2 = output tag
set -x
mkdir -p gpurun_out
PROBE_QUICK=1 timeout 400 ncu --target-processes all --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$1" -c 24 --csv --log-file gpurun_out/ncu_nvlink_$2.csv python scripts/peer_traffic_probe.py 2 > gpurun_out/ncu_nvlink_$2.log 2>&1; echo rc $?
tail -3 gpurun_out/ncu_nvlink_$2.log
