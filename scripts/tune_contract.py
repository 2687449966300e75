"""Kernel-tuning probe: the contraction kernel on the shapes that make up the CCSD step (and the raw GEMM view), for the
library named by SIPGPU_LIB (default: the in-tree build).  Prints one JSON line.  Development tool, not a bench."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip  # noqa: E402

api = sip.api
sip.init(0)
torch.cuda.set_device(0)
stream = torch.cuda.ExternalStream(api.stream_handle())
out = {"lib": os.path.basename(sip.lib_path())}


def time_ms(fn, reps=3, warm=1):
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


out["dmma_peak"] = round(max(api.dmma_peak_probe(40000) for _ in range(2)), 2)
for n in (4096, 8192):
    A, B, Cc = api.DeviceBlock((n, n)).fill(0.5), api.DeviceBlock((n, n)).fill(0.25), api.DeviceBlock((n, n))
    ms = time_ms(lambda: api.dgemm_tn(n, n, n, A, n, B, n, Cc, n), reps=4)
    out[f"gemm{n}"] = round(2.0 * n ** 3 / ms / 1e9, 2)
    del A, B, Cc

v, o = 50, 20
T = [api.DeviceBlock((v, o, v, o)).fill(0.25) for _ in range(16)]
V = [api.DeviceBlock((v, o, v, o)).fill(0.5) for _ in range(16)]


def chained(name, dl, ll, rl, lsh, rsh, Ls, Rs, nd, cl, flops_pair):
    D = [api.DeviceBlock((v, o, v, o)) for _ in range(nd)]
    ptrn, ierr = api.get_contraction_ptrn(dl, ll, rl)
    assert ierr == 0
    bc = api.BatchedContraction(ptrn, [lsh] * nd, [rsh] * nd, [(v, o, v, o)] * nd,
                                [Ls[(i * 5) % len(Ls)].ptr for i in range(nd * cl)],
                                [Rs[(i * 3) % len(Rs)].ptr for i in range(nd * cl)], [d.ptr for d in D],
                                chain_start=[i * cl for i in range(nd + 1)])
    ms = time_ms(lambda: bc.launch(), reps=3, warm=1)
    out[name] = round(nd * cl * flops_pair / ms / 1e9, 2)


# the five terms of the step, one destination-stationary work-list each (labels as in sial_workload.TERMS)
num = {c: i + 1 for i, c in enumerate("abcdijkl")}
L = lambda s: [num[c] for c in s]
chained("ring1", L("aibj"), L("aick"), L("ckbj"), (v, o, v, o), (v, o, v, o), V, T, 148, 12 * 3, 2e9)
chained("ring3", L("aibj"), L("akcj"), L("bcki"), (v, o, v, o), (v, v, o, o), T, [api.DeviceBlock((v, v, o, o)).fill(0.5) for _ in range(8)], 148, 12 * 3, 2e9)
chained("hh", L("aibj"), L("akbl"), L("ikjl"), (v, o, v, o), (o, o, o, o), T, [api.DeviceBlock((o, o, o, o)).fill(0.5) for _ in range(9)], 296, 9, 2.0 * 2500 * 400 * 400)
ao = [api.DeviceBlock((v, v, v, v)).fill(0.5) for _ in range(4)]
chained("pp", L("aibj"), L("cadb"), L("cidj"), (v, v, v, v), (v, o, v, o), ao, T, 148, 36, 2.0 * 2500 * 400 * 2500)
print(json.dumps(out), flush=True)
