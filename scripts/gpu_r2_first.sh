#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget ran out, then the round-end sequence.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_first.sh'
set -u
mkdir -p gpurun_out
# 1. the whole GPU suite (includes tests/test_gpu_z_lccd_water_energy.py = the reference's LCCD energy golden on the device,
#    never run on a GPU yet; it sorts last and prints the energies with -s; also the level-1 composition of tests/test_gpu_ref_block_on_sipgpu.py, never run on a GPU yet,
#    and everything behind the export map / SIPGPU_NO_TENSORDIL_PROTOTYPES changes)
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu_r2_first.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_r2_first.log
# 2. smoke + the CPU arm with the fair thread policy (expected near 0.19 TFLOP/s on 16 cores, was 0.119)
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r2.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; echo "ref arm rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_ref_r2.json')); print('cpu arm', d['value'], 'TFLOP/s on', d['cpu_baseline']['cores'], 'cores')"
# 3. the default bench line
timeout 900 python bench.py > gpurun_out/bench_n1_r2.json 2> gpurun_out/bench_n1_r2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_n1_r2.json
# 4. the energy tests with their printed energies, and the real CC programs op-at-a-time vs recorded (tiny molecules: launch / host bound)
timeout 600 python -m pytest tests/test_gpu_z_lccd_water_energy.py tests/test_gpu_z_cross_product.py -m gpu -q -s > gpurun_out/pytest_gpu_energy_r2.log 2>&1; echo "energy tests rc=$?"; grep -a "on the device\|patterns at\|passed\|failed" gpurun_out/pytest_gpu_energy_r2.log | tail -20
timeout 600 python scripts/cc_programs_bench.py > gpurun_out/cc_programs_bench.jsonl 2> gpurun_out/cc_programs_bench.err; echo "cc bench rc=$?"; cut -c1-200 gpurun_out/cc_programs_bench.jsonl
# 5. (with `gpurun --gpus 2`) the energy goldens on 2 GPUs:
#    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_n_energy.py
