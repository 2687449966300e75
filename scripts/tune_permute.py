"""Permute-kernel probe (development): GB/s (16 bytes per element) of single and batched permutes."""
import itertools, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip
api = sip.api
sip.init(0)
stream = torch.cuda.ExternalStream(api.stream_handle())


def time_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sip.sync()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"copy_gbs": round(api.copy_bw_probe(1 << 30, 10), 1)}
for shape in ((64, 64, 64, 64), (32, 32, 32, 32), (50, 20, 50, 20)):
    a, b = api.DeviceBlock(shape).fill(1.0), api.DeviceBlock(shape)
    res = {}
    for p in itertools.permutations(range(4)):
        if p == (0, 1, 2, 3):
            continue
        transp = [1] + [x + 1 for x in p]
        res["".join(map(str, p))] = 16.0 * np.prod(shape) / time_ms(lambda: api.permute(a, transp, out=b)) / 1e6
    v = sorted(res.values())
    out[f"single{shape}"] = {"min": round(v[0]), "median": round(v[len(v) // 2]), "max": round(v[-1])}
# batched: 128 blocks of (50,20,50,20) per launch (1 GiB of traffic per launch... 128 x 16 MB = 2 GB)
shape = (50, 20, 50, 20)
n = 128
ins = [api.DeviceBlock(shape).fill(1.0) for _ in range(n)]
outs = [api.DeviceBlock(shape) for _ in range(n)]
res = {}
for p in ((0, 3, 2, 1), (2, 3, 0, 1), (1, 0, 3, 2), (3, 2, 1, 0), (2, 1, 0, 3)):
    transp = [1] + [x + 1 for x in p]
    res["".join(map(str, p))] = round(n * 16.0 * np.prod(shape) / time_ms(lambda: api.permute_batched(ins, transp, outs), reps=5) / 1e6)
    res["".join(map(str, p)) + "+acc"] = round(n * 24.0 * np.prod(shape) / time_ms(lambda: api.permute_batched(ins, transp, outs, alpha=0.5, beta=1.0), reps=5) / 1e6)
out["batched128x(50,20,50,20)"] = res
print(json.dumps(out), flush=True)
