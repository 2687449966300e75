#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for ls in 0 1; do
  SIPGPU_LOCKSTEP=$ls timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_lockstep$ls.json 2> gpurun_out/bench_lockstep$ls.err
  echo "lockstep=$ls rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_lockstep$ls.json'));print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])"
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:contract_kernel --csv \
  --log-file gpurun_out/traffic_ccsd_full.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_traffic.log 2>&1
echo "ncu traffic rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/traffic_ccsd_full.csv')))
hdr=None; data={}
for r in rows:
    if r and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); k=int(d["ID"])
        data.setdefault(k,{"name":d["Kernel Name"]})[d["Metric Name"]]=float(d["Metric Value"].replace(",",""))
for k,v in sorted(data.items()):
    t=v["gpu__time_duration.sum"]/1e9
    print(k, v["name"][-34:], f"{t:.3f}s read {v['dram__bytes_read.sum']/1e9:.1f} GB write {v['dram__bytes_write.sum']/1e9:.1f} GB")
PY
