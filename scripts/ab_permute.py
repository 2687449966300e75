"""A/B helper (development): batched permute GB/s, all 23 rank-4 permutations, for the library named by SIPGPU_LIB."""
import itertools, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip
api = sip.api
sip.init(0)
stream = torch.cuda.ExternalStream(api.stream_handle())


def time_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


api.set_tuning("permute_bulk", int(os.environ.get("PERMUTE_BULK", "3")))
api.set_tuning("permute_vec", int(os.environ.get("PERMUTE_VEC", "-1")))
out = {}
for shape, n in (((16,) * 4, 1024), ((32,) * 4, 128), ((64,) * 4, 8), ((50, 20, 50, 20), 128)):
    ins = [api.DeviceBlock(shape).fill(1.0) for _ in range(n)]
    outs = [api.DeviceBlock(shape) for _ in range(n)]
    res, acc, per = [], [], {}
    for p in itertools.permutations(range(4)):
        if p == (0, 1, 2, 3):
            continue
        transp = [1] + [x + 1 for x in p]
        bp = api.BatchedPermute(ins, transp, outs)   # pointer arrays marshalled once: time the library, not ctypes
        res.append(n * 16.0 * np.prod(shape) / time_ms(lambda: bp.launch()) / 1e6)
        acc.append(n * 24.0 * np.prod(shape) / time_ms(lambda: bp.launch(alpha=0.5, beta=1.0)) / 1e6)
        per["".join(map(str, p))] = [round(res[-1]), round(acc[-1])]
    out[str(shape)] = {"blocks": n, "min": round(min(res)), "median": round(float(np.median(res))), "max": round(max(res)),
                       "acc_median": round(float(np.median(acc))), "per_perm": per}
    del ins, outs
print(json.dumps({"permute_bulk": int(os.environ.get("PERMUTE_BULK", "3")), "permute_vec": int(os.environ.get("PERMUTE_VEC", "-1")), "GBps": out}), flush=True)
