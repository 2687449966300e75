#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
