#!/bin/bash
# final check of round 1 after the host-side work on the deferred op stream (arena allocation, memoised validation,
# persistent op buffer): the whole GPU suite, then the C++ pardo driver at three segmentations
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r16.log
: > gpurun_out/wl_pardo_bench_r16.jsonl
for cfg in "16 3 16 6" "8 4 24 6" "20 3 50 6"; do
  timeout 25 scripts/micro/wl_pardo_bench $cfg 2 >> gpurun_out/wl_pardo_bench_r16.jsonl 2>> gpurun_out/wl_pardo_bench_r16.err; echo "cfg $cfg rc=$?"
done
python - <<'PY'
import json
for line in open('gpurun_out/wl_pardo_bench_r16.jsonl'):
    d = json.loads(line)
    print(d['workload'][-22:], 'eager', d['op_at_a_time']['seconds'], 'recorded', d['recorded']['seconds'], d['recorded']['tflops'], 'TF/s speedup', d['speedup'], 'diff', d['t2new_norm2_rel_diff'])
PY
tail -2 gpurun_out/wl_pardo_bench_r16.err
