#!/bin/bash
# N = 2: peer traffic probe (TMA bulk vs LDG/red.add), the two-process distributed test, the bench line with the nvlink key
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 300 python scripts/peer_traffic_probe.py 1 > gpurun_out/peer_traffic_w1.log 2>&1; echo rc $?
timeout 300 python scripts/peer_traffic_probe.py 2 > gpurun_out/peer_traffic_w2.log 2>&1; echo rc $?
tail -30 gpurun_out/peer_traffic_w2.log
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_worklist.py -x -q 2>&1 | tail -5 > gpurun_out/multi_test.txt
cat gpurun_out/multi_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --wall-budget 200 > gpurun_out/bench_n2_traffic.json 2> gpurun_out/bench_n2_traffic.err
tail -2 gpurun_out/bench_n2_traffic.json
