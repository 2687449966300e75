"""A handful of low-intensity contraction shapes through both kernels (bandwidth-shaped lowint.cu vs the 128-wide tile
kernel, switched with sipgpu_set_tuning), timed with CUDA events: ms, algorithmic GB/s.  For ncu: --only <name> --route lowint"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces4_b200 as sip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default=None)
ap.add_argument("--route", default=None)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
api = sip.api
sip.init(0)
torch.cuda.set_device(0)
stream = torch.cuda.ExternalStream(api.stream_handle())
CASES = {
    "gemv_mfast": ("ab", "abcd", "cd", dict(a=20, b=50, c=50, d=20), 1024),
    "gemv_kfast": ("ab", "cdab", "dc", dict(a=50, b=20, c=20, d=50), 256),
    "skinny": ("abcd", "abce", "ed", dict(a=20, b=50, c=20, d=50, e=50), 100),
    "skinny_T": ("abcd", "ecba", "ed", dict(a=20, b=50, c=50, d=50, e=20), 100),
    "rank2_k": ("ab", "cade", "cbde", dict(a=50, b=20, c=50, d=50, e=20), 100),
    "rank2_m": ("ab", "acde", "bcde", dict(a=50, b=50, c=20, d=50, e=20), 100),
    "tiny": ("ab", "ac", "cb", dict(a=20, b=20, c=50), 4096),
    "dot": ("ab", "cda", "cdb", dict(a=1, b=1, c=50, d=20), 4096),
}
for name, (d, l, r, ext, nb) in CASES.items():
    if args.only and name != args.only:
        continue
    labs = sorted(set(d + l + r))
    num = {c: i + 1 for i, c in enumerate(labs)}
    lsh, rsh, dsh = [ext[c] for c in l], [ext[c] for c in r], [ext[c] for c in d]
    ptrn, ierr = api.get_contraction_ptrn([num[c] for c in d], [num[c] for c in l], [num[c] for c in r])
    assert ierr == 0
    K = float(np.prod([ext[c] for c in l if c in r]))
    flops = 2.0 * np.prod(dsh) * K * nb
    byts = 8.0 * (np.prod(lsh) + np.prod(rsh) + np.prod(dsh)) * nb

    def pool(shape, tag):
        n = int(max(1, min(nb, 1.5e9 // (8 * np.prod(shape)))))
        return [api.DeviceBlock(shape).fill_hash(tag, i, 1.0) for i in range(n)]
    Ls, Rs, Ds = pool(lsh, 1), pool(rsh, 2), pool(dsh, 3)
    bc = api.BatchedContraction(ptrn, [lsh] * nb, [rsh] * nb, [dsh] * nb, [Ls[i % len(Ls)].ptr for i in range(nb)],
                                [Rs[i % len(Rs)].ptr for i in range(nb)], [Ds[i % len(Ds)].ptr for i in range(nb)])
    out = {"case": name, "blocks": nb, "MB": byts / 1e6}
    for route in ("lowint", "tiles"):
        if args.route and route != args.route:
            continue
        api.set_tuning("lowint_scope", 2 if route == "lowint" else 0)
        bc.launch()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            bc.launch()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        out[route] = {"ms": round(ms, 4), "GBps": round(byts / ms / 1e6), "TFs": round(flops / ms / 1e9, 2)}
    print(json.dumps(out), flush=True)
    del Ls, Rs, Ds, bc
