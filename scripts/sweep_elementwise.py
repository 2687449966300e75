"""HBM-bound block ops (elementwise.cu, worklist.cu) against the measured copy bandwidth: fill 8n, scale 16n,
scale_and_copy 16n, axpy 24n, add_sub 24n, dot 16n bytes per element, for one 8 MB block per launch, one 128 MB block
per launch, and 256 blocks of 512 KB issued one by one vs recorded into the deferred op stream (one batched launch).
Writes gpurun_out/sweep_elementwise.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces4_b200 as sip  # noqa: E402

api = sip.api
sip.init(0)
torch.cuda.set_device(0)
stream = torch.cuda.ExternalStream(api.stream_handle())
flush_buf = api.DeviceBlock((48 * 1024 * 1024,))


def time_ms(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush_buf.fill(0.0)   # 384 MB > L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


bw = api.copy_bw_probe(1 << 30, 10)
out = {"copy_gbs": bw, "single": {}, "batched_256x512KB": {}}
for n in (1 << 20, 1 << 24):
    a, b, c = api.DeviceBlock((n,)).fill(1.0), api.DeviceBlock((n,)).fill(2.0), api.DeviceBlock((n,)).fill(3.0)
    acc = api.DeviceBlock((1,), zero=True)
    ops = {"fill": (8, lambda: a.fill(0.5)), "scale": (16, lambda: a.scale(1.0001)),
           "scale_and_copy": (16, lambda: a.scale_and_copy(b, 0.5)), "axpy": (24, lambda: a.axpy(b, 0.5)),
           "add_sub": (24, lambda: a.set_add_sub(b, c, 1.0)),
           "dot_accumulate": (16, lambda: api._check(api.lib().sipgpu_block_dot_accumulate(a.ptr, b.ptr, n, acc.ptr)))}
    row = {}
    for name, (bpe, fn) in ops.items():
        ms = time_ms(fn)
        row[name] = {"ms": ms, "gbs": bpe * n / ms / 1e6, "frac_of_copy_peak": bpe * n / ms / 1e6 / bw}
    out["single"][f"{8 * n >> 20}MB"] = row
    print(f"{8 * n >> 20} MB block: " + ", ".join(f"{k} {v['gbs']:.0f} GB/s" for k, v in row.items()), flush=True)
    del a, b, c

nb, n = 256, 1 << 16
A = [api.DeviceBlock((n,)).fill(1.0) for _ in range(nb)]
B = [api.DeviceBlock((n,)).fill(2.0) for _ in range(nb)]


def eager():
    for x, y in zip(A, B):
        x.axpy(y, 0.5)


def recorded():
    api.wl_begin()
    for x, y in zip(A, B):
        x.axpy(y, 0.5)
    api.wl_end()


for name, fn in (("op_at_a_time", eager), ("recorded", recorded)):
    ms = time_ms(fn, reps=5)
    out["batched_256x512KB"][name] = {"ms": ms, "gbs": 24.0 * n * nb / ms / 1e6}
    print(f"256 x 512 KB axpy {name}: {ms:.3f} ms, {24.0 * n * nb / ms / 1e6:.0f} GB/s", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep_elementwise.json"), "w"), indent=1)
