"""Level-0 integration: the libtensordil host-pointer ABI (tensor_block_contract__ / tensor_block_copy__)
exactly as an unmodified aces4 would call it -- pageable host buffers in, host buffers out, every call staged through the
device -- next to the CPU oracle (reference algorithm, OpenBLAS dgemm, all host cores) on the same blocks.
Writes gpurun_out/level0_bench.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ may use the oracle (test infrastructure)
sys.path.insert(0, ROOT)
import aces4_b200 as sip  # noqa: E402
from oracle import oracle  # noqa: E402  (the CPU arm of this comparison)

api = sip.api
sip.init(0)
cores = len(os.sched_getaffinity(0))
blas = oracle.use_openblas(cores)
rng = np.random.default_rng(1)
out = {"cores": cores, "blas": blas.split()[0], "cases": []}
for name, dl, ll, rl, ext in (
        ("ring  D[a,i,b,j]=L[a,i,c,k]*R[c,k,b,j] v=50 o=20", "aibj", "aick", "ckbj", dict(a=50, b=50, c=50, i=20, j=20, k=20)),
        ("hh    D[a,i,b,j]=L[a,k,b,l]*R[i,k,j,l] v=50 o=20", "aibj", "akbl", "ikjl", dict(a=50, b=50, i=20, j=20, k=20, l=20)),
        ("skinny D[a,i,b,j]=L[a,i,c,j]*R[c,b]    v=50 o=20", "aibj", "aicj", "cb", dict(a=50, b=50, c=50, i=20, j=20)),
        ("small D[a,i,b,j]=L[a,i,c,k]*R[c,k,b,j] seg 16", "aibj", "aick", "ckbj", dict(a=16, b=16, c=16, i=16, j=16, k=16))):
    labs = sorted(ext)
    num = {c: i + 1 for i, c in enumerate(labs)}
    L = np.asfortranarray(rng.uniform(-1, 1, [ext[c] for c in ll]))
    R = np.asfortranarray(rng.uniform(-1, 1, [ext[c] for c in rl]))
    dext = [ext[c] for c in dl]
    ptrn, ierr = api.get_contraction_ptrn([num[c] for c in dl], [num[c] for c in ll], [num[c] for c in rl])
    flops = 2.0 * np.prod(dext) * np.prod([ext[c] for c in ll if c in rl])

    def gpu():
        d, ie = api.tensor_block_contract(ptrn, L, R, dext)
        assert ie == 0
        return d

    def cpu():
        d, ie = oracle.block_contract(ptrn, L, R, dext)
        assert ie == 0
        return d

    res = {}
    for tag, fn, reps in (("gpu_level0", gpu, 10), ("cpu_oracle", cpu, 3)):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            d = fn()
        res[tag] = (time.perf_counter() - t0) / reps
        res[tag + "_result"] = d
    err = np.max(np.abs(res["gpu_level0_result"] - res["cpu_oracle_result"])) / np.max(np.abs(res["cpu_oracle_result"]))
    row = {"case": name, "gflop": flops / 1e9, "bytes_staged": 8.0 * (L.size + R.size + np.prod(dext)),
           "gpu_level0_ms": res["gpu_level0"] * 1e3, "cpu_oracle_ms": res["cpu_oracle"] * 1e3,
           "gpu_level0_tflops": flops / res["gpu_level0"] / 1e12, "cpu_tflops": flops / res["cpu_oracle"] / 1e12,
           "speedup": res["cpu_oracle"] / res["gpu_level0"], "rel_err": float(err)}
    out["cases"].append(row)
    print(f"{name}: level-0 {row['gpu_level0_ms']:.2f} ms ({row['gpu_level0_tflops']:.2f} TF/s) vs CPU {row['cpu_oracle_ms']:.2f} ms "
          f"({row['cpu_tflops']:.3f} TF/s): {row['speedup']:.1f}x, rel.err {err:.1e}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "level0_bench.json"), "w"), indent=1)
