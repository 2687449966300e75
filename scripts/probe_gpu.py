"""First-contact probe on a B200: FP64 tensor (DMMA) issue-rate peak, copy bandwidth, the fused contraction kernel at
GEMM-like and permuted patterns, cuBLAS DGEMM (torch.matmul fp64) beside it, and the permute kernel.
Writes gpurun_out/probe.json.  Not a bench: exploratory numbers for DESIGN.md / kernel tuning."""
import itertools
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip  # noqa: E402

api = sip.api
sip.init(0)
torch.cuda.set_device(0)
ext_stream = torch.cuda.ExternalStream(api.stream_handle())
out = {}


def time_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sip.sync()
    e0.record(ext_stream)
    for _ in range(reps):
        fn()
    e1.record(ext_stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


out["dmma_peak_tflops"] = [api.dmma_peak_probe(it) for it in (20000, 50000, 50000)]
out["copy_gbs"] = [api.copy_bw_probe(1 << 30, 10), api.copy_bw_probe(4 << 30, 5)]
print("dmma peak TF/s", out["dmma_peak_tflops"], "copy GB/s", out["copy_gbs"], flush=True)

# cuBLAS DGEMM through torch (library bar)
cub = {}
for n in (1024, 2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a.t(), b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 20 if n <= 2048 else 5
    e0.record()
    for _ in range(reps):
        torch.matmul(a.t(), b)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    cub[n] = 2.0 * n ** 3 / ms / 1e9
    del a, b
out["cublas_dgemm_tn_tflops"] = cub
print("cuBLAS dgemm TN TF/s", cub, flush=True)

# our kernel, raw GEMM view
ours = {}
for n in (256, 1024, 2048, 4096, 8192):
    A, B, Cc = api.DeviceBlock((n, n)), api.DeviceBlock((n, n)), api.DeviceBlock((n, n))
    A.fill(0.5), B.fill(0.25)
    ms = time_ms(lambda: api.dgemm_tn(n, n, n, A, n, B, n, Cc, n), reps=20 if n <= 2048 else 5)
    ours[n] = 2.0 * n ** 3 / ms / 1e9
    del A, B, Cc
out["sipgpu_dgemm_tn_tflops"] = ours
print("sipgpu dgemm_tn TF/s", ours, flush=True)

# contraction patterns at the CCSD block shape (v=50, o=20) and at uniform s
pat = {}
cases = {
    "ring  Z[a,i,b,j]=T[a,i,c,k]*V[c,k,b,j]": ([1, 2, 3, 4], [1, 2, 5, 6], [5, 6, 3, 4]),
    "ring2 Z[a,i,b,j]=T[a,k,c,i]*V[c,k,b,j]": ([1, 2, 3, 4], [1, 6, 5, 2], [5, 6, 3, 4]),
    "ring3 Z[a,i,b,j]=T[b,k,c,i]*V[a,k,c,j]": ([1, 2, 3, 4], [3, 6, 5, 2], [1, 6, 5, 4]),
    "hh    Z[a,i,b,j]=W[i,k,j,l]*T[a,k,b,l]": ([1, 2, 3, 4], [2, 5, 4, 6], [1, 5, 3, 6]),
    "aolad Y[m,i,n,j]=V[l,m,s,n]*T[l,i,s,j]": ([1, 2, 3, 4], [5, 1, 6, 3], [5, 2, 6, 4]),
}
for name, (dl, ll, rl) in cases.items():
    for (v, o) in ((50, 20), (32, 32), (64, 64)):
        kind = {1: v, 3: v, 5: v, 2: o, 4: o, 6: o}
        if name.startswith("hh"):
            kind = {1: v, 3: v, 2: o, 4: o, 5: o, 6: o}
        if name.startswith("aolad"):
            kind = {1: v, 3: v, 5: v, 6: v, 2: o, 4: o}
        lext, rext, dext = [kind[x] for x in ll], [kind[x] for x in rl], [kind[x] for x in dl]
        if np.prod(lext) > 2 ** 30 or np.prod(rext) > 2 ** 30:
            continue
        Lb, Rb, Db = api.DeviceBlock(lext), api.DeviceBlock(rext), api.DeviceBlock(dext)
        Lb.fill(0.5), Rb.fill(0.25)
        flops = 2.0 * np.prod(dext) * np.prod([kind[x] for x in ll if x not in dl])
        ms = time_ms(lambda: api.contract_labels(dl, dext, ll, Lb, rl, Rb, out=Db), reps=5)
        pat[f"{name} v={v} o={o}"] = {"ms": ms, "tflops": flops / ms / 1e9}
        print(name, v, o, f"{ms:.3f} ms {flops / ms / 1e9:.2f} TF/s", flush=True)
        del Lb, Rb, Db
out["contract_patterns"] = pat

# batched: 1296-block-class work-list of ring terms at (50,20,50,20): 256 problems
n = 256
Ls = [api.DeviceBlock((50, 20, 50, 20)).fill(0.5) for _ in range(16)]
Rs = [api.DeviceBlock((50, 20, 50, 20)).fill(0.25) for _ in range(16)]
Ds = [api.DeviceBlock((50, 20, 50, 20)) for _ in range(n)]
ptrn, _ = api.get_contraction_ptrn([1, 2, 3, 4], [1, 2, 5, 6], [5, 6, 3, 4])
bc = api.BatchedContraction(ptrn, [(50, 20, 50, 20)] * n, [(50, 20, 50, 20)] * n, [(50, 20, 50, 20)] * n,
                            [Ls[i % 16].ptr for i in range(n)], [Rs[(i * 7) % 16].ptr for i in range(n)],
                            [d.ptr for d in Ds])
ms = time_ms(lambda: bc.launch(), reps=3, warm=1)
out["batched_ring_256x(1000^3)"] = {"ms": ms, "tflops": n * 2e9 / ms / 1e9}
print("batched ring 256 blocks", out["batched_ring_256x(1000^3)"], flush=True)
# chained: 148 destination blocks, each the sum of 12 operand pairs (ring term over 12 contracted segments)
nd, cl = 148, 12
Ds2 = Ds[:nd]
chain = [i * cl for i in range(nd + 1)]
bc2 = api.BatchedContraction(ptrn, [(50, 20, 50, 20)] * nd, [(50, 20, 50, 20)] * nd, [(50, 20, 50, 20)] * nd,
                             [Ls[(i * 5) % 16].ptr for i in range(nd * cl)], [Rs[(i * 3) % 16].ptr for i in range(nd * cl)],
                             [d.ptr for d in Ds2], chain_start=chain)
ms = time_ms(lambda: bc2.launch(), reps=3, warm=1)
out["chained_ring_148x12x(1000^3)"] = {"ms": ms, "tflops": nd * cl * 2e9 / ms / 1e9}
print("chained ring 148 dest x 12 pairs", out["chained_ring_148x12x(1000^3)"], flush=True)
del Ls, Rs, Ds, Ds2
# pp-ladder stand-in: Y[mu,i,nu,j] = sum_{lambda,sigma segs} aoint[lambda,mu,sigma,nu] * T[lambda,i,sigma,j]
# (M = 2500, N = 400, K = 2500 per pair): 74 destinations x 6 pairs
nd, cl = 74, 6
ao = [api.DeviceBlock((50, 50, 50, 50)).fill(0.5) for _ in range(4)]
tt = [api.DeviceBlock((50, 20, 50, 20)).fill(0.25) for _ in range(8)]
yy = [api.DeviceBlock((50, 20, 50, 20)) for _ in range(nd)]
ptrn_pp, _ = api.get_contraction_ptrn([1, 2, 3, 4], [5, 1, 6, 3], [5, 2, 6, 4])
bc3 = api.BatchedContraction(ptrn_pp, [(50, 50, 50, 50)] * nd, [(50, 20, 50, 20)] * nd, [(50, 20, 50, 20)] * nd,
                             [ao[(i * 3) % 4].ptr for i in range(nd * cl)], [tt[(i * 5) % 8].ptr for i in range(nd * cl)],
                             [d.ptr for d in yy], chain_start=[i * cl for i in range(nd + 1)])
ms = time_ms(lambda: bc3.launch(), reps=3, warm=1)
out["chained_ppladder_74x6x(2500x400x2500)"] = {"ms": ms, "tflops": nd * cl * 2.0 * 2500 * 400 * 2500 / ms / 1e9}
print("chained pp ladder", out["chained_ppladder_74x6x(2500x400x2500)"], flush=True)
# hh ladder: M = 2500, N = 400, K = 400 per pair, 9 pairs
nd, cl = 296, 9
vo = [api.DeviceBlock((20, 20, 20, 20)).fill(0.5) for _ in range(9)]
ptrn_hh, _ = api.get_contraction_ptrn([1, 2, 3, 4], [1, 5, 3, 6], [2, 5, 4, 6])
yy2 = [api.DeviceBlock((50, 20, 50, 20)) for _ in range(nd)]
bc4 = api.BatchedContraction(ptrn_hh, [(50, 20, 50, 20)] * nd, [(20, 20, 20, 20)] * nd, [(50, 20, 50, 20)] * nd,
                             [tt[(i * 5) % 8].ptr for i in range(nd * cl)], [vo[i % 9].ptr for i in range(nd * cl)],
                             [d.ptr for d in yy2], chain_start=[i * cl for i in range(nd + 1)])
ms = time_ms(lambda: bc4.launch(), reps=3, warm=1)
out["chained_hhladder_296x9x(2500x400x400)"] = {"ms": ms, "tflops": nd * cl * 2.0 * 2500 * 400 * 400 / ms / 1e9}
print("chained hh ladder", out["chained_hhladder_296x9x(2500x400x400)"], flush=True)
del ao, tt, yy, yy2, vo

# permutes: all 24 patterns at (50,20,50,20) and 32^4, 64^4
perm = {}
for shape in ((50, 20, 50, 20), (32, 32, 32, 32), (64, 64, 64, 64)):
    a = api.DeviceBlock(shape).fill(1.0)
    b = api.DeviceBlock(shape)
    nbytes = 16.0 * np.prod(shape)
    res = {}
    for p in itertools.permutations(range(4)):
        transp = [1] + [x + 1 for x in p]
        ms = time_ms(lambda: api.permute(a, transp, out=b), reps=10)
        res["".join(map(str, p))] = nbytes / ms / 1e6
    perm[str(shape)] = res
    vals = list(res.values())
    print("permute", shape, "GB/s min/median/max", min(vals), float(np.median(vals)), max(vals), flush=True)
    del a, b
out["permute_gbs"] = perm

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print("launches", sip.kernel_launches())
