"""Config 4 of BASELINE.json (single-GPU block-contraction sweep) plus the (T)- and EOM-shaped micro-workloads
(SURVEY.md 8d items 2, 4, 5).  Writes gpurun_out/sweep.json.  Not the headline bench (bench.py is); these are the
per-pattern / per-size numbers behind it.

  * rank-4 contractions with two contracted indices at segment size s in {16,24,32,40,48,64} and one ragged case, over
    the (4,4,4) label patterns that occur in the reference's SIAL programs (tests/golden/sial_contraction_patterns.txt);
    each pattern is timed as ONE block per launch and as a work-list of nb blocks per launch (nb chosen so that a
    launch holds >= 30 GFLOP) -- the SIAL regime is thousands of small blocks per pardo;
  * pure permutes, all 24 rank-4 patterns, every s, one block and batched;
  * (T) inner contraction  tpppp[a,a1,b,k1] = t1ppp[b1,a1,b] * t2ppp[b1,a,k1]  accumulated into slices of a rank-6
    block (rccsdpt_aab.sialx:526-580): fused (beta = 1 into the slice) vs unfused (temp block + accumulate);
  * EOM-style rank-5 blocks with an extent-1 leading index contracted with rank-2 arrays.
"""
import itertools
import json
import os
import sys

# CPU column: the oracle's OpenMP loops and OpenBLAS's threads take turns; idle libgomp workers must not spin through the
# dgemm (see bench.py) -- set before any OpenMP runtime is loaded
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("GOMP_SPINCOUNT", "0")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aces4_b200 as sip  # noqa: E402
from conftest import sial_patterns  # noqa: E402

api = sip.api
sip.init(0)
torch.cuda.set_device(0)
stream = torch.cuda.ExternalStream(api.stream_handle())
quick = "--quick" in sys.argv
flush_buf = api.DeviceBlock((48 * 1024 * 1024,))  # 384 MB > 126 MB L2


def time_ms(fn, reps=5, warm=2, flush=False):
    for _ in range(warm):
        fn()
    times = []
    for _ in range(reps):
        if flush:
            flush_buf.fill(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times))


out = {"dmma_peak_tflops": max(api.dmma_peak_probe(40000) for _ in range(2)), "copy_gbs": api.copy_bw_probe(1 << 30, 10)}
peak = out["dmma_peak_tflops"]
pats444 = [(d, l, r) for d, l, r, kinds, where in sial_patterns() if (len(d), len(l), len(r)) == (4, 4, 4)]
sizes = [16, 32, 64] if quick else [16, 24, 32, 40, 48, 64]

# ---------------------------------------------------------------------------------------------- contractions
con = {}
for s in sizes + ["ragged"]:
    rows = []
    for d, l, r in pats444[:6] if quick else pats444:
        labs = sorted(set(d + l + r))
        num = {c: i + 1 for i, c in enumerate(labs)}
        if s == "ragged":
            exts = dict(zip(labs, itertools.cycle([13, 30, 50, 64, 30, 13])))
        else:
            exts = {c: s for c in labs}
        lsh, rsh, dsh = [exts[c] for c in l], [exts[c] for c in r], [exts[c] for c in d]
        flops = 2.0 * np.prod(dsh) * np.prod([exts[c] for c in l if c in r])
        ptrn, ierr = api.get_contraction_ptrn([num[c] for c in d], [num[c] for c in l], [num[c] for c in r])
        assert ierr == 0
        nb = int(max(1, min(2048, np.ceil(30e9 / flops))))
        nb = 37 * int(np.ceil(nb / 37))  # whole waves of tiles on 148 SMs (4 | tiles per block at uniform s)
        pool = 8
        Ls = [api.DeviceBlock(lsh).fill_hash(1, i, 1.0) for i in range(pool)]
        Rs = [api.DeviceBlock(rsh).fill_hash(2, i, 1.0) for i in range(pool)]
        nD = min(nb, max(1, int(2e9 // (8 * np.prod(dsh)))))
        Ds = [api.DeviceBlock(dsh) for _ in range(nD)]
        ms1 = time_ms(lambda: api.contract(ptrn, Ls[0], Rs[0], dsh, out=Ds[0]), reps=7)
        ms1c = time_ms(lambda: api.contract(ptrn, Ls[0], Rs[0], dsh, out=Ds[0]), reps=5, flush=True)
        bc = api.BatchedContraction(ptrn, [lsh] * nb, [rsh] * nb, [dsh] * nb, [Ls[i % pool].ptr for i in range(nb)],
                                    [Rs[(3 * i) % pool].ptr for i in range(nb)], [Ds[i % nD].ptr for i in range(nb)])
        msb = time_ms(lambda: bc.launch(), reps=3, warm=1)
        rows.append({"pattern": f"{d}={l}*{r}", "gflop_per_block": flops / 1e9, "one_block_ms_warm": ms1,
                     "one_block_ms_cold": ms1c, "one_block_tflops": flops / ms1 / 1e9, "batch": nb, "batch_ms": msb,
                     "batch_tflops": nb * flops / msb / 1e9, "batch_frac_of_dmma_peak": nb * flops / msb / 1e9 / peak})
        del Ls, Rs, Ds, bc
    con[str(s)] = rows
    v = [x["batch_tflops"] for x in rows]
    v1 = [x["one_block_tflops"] for x in rows]
    print(f"contract s={s}: batched TF/s min/median/max {min(v):.1f}/{np.median(v):.1f}/{max(v):.1f}; one block "
          f"{min(v1):.2f}/{np.median(v1):.2f}/{max(v1):.2f}", flush=True)
out["contract_444"] = con

# ------------------------------------------------------------------------ CPU column of config 4 ("1 GPU vs CPU")
# The reference's CPU path for the same block contraction -- permute L, permute R, dgemm('T','N'), permute D
# (tensor_dil_omp.F90:662-796), restated by the oracle with OpenBLAS dgemm -- on this box's host cores: at 8 threads (the
# reference caps OpenMP at 8, src/sip/core/sip.cpp:113) and at all cores; one block per call, as the interpreter issues them.
if "--no-cpu" not in sys.argv:
    import ctypes
    import time

    from oracle import oracle

    cores = len(os.sched_getaffinity(0))
    cpu = {}
    for s_ in sizes + ["ragged"]:
        per_threads = {}
        for nthr in sorted({min(8, cores), cores}):
            oracle.use_openblas(nthr)
            try:
                ctypes.CDLL("libgomp.so.1").omp_set_num_threads(nthr)
            except OSError:
                pass
            tf = []
            for d, l, r in pats444[:4]:
                labs = sorted(set(d + l + r))
                num = {c: i + 1 for i, c in enumerate(labs)}
                exts = dict(zip(labs, itertools.cycle([13, 30, 50, 64, 30, 13]))) if s_ == "ragged" else {c: s_ for c in labs}
                rng = np.random.default_rng(7)
                Lh = np.asfortranarray(rng.uniform(-1, 1, [exts[c] for c in l]))
                Rh = np.asfortranarray(rng.uniform(-1, 1, [exts[c] for c in r]))
                dsh = [exts[c] for c in d]
                flops = 2.0 * np.prod(dsh) * np.prod([exts[c] for c in l if c in r])
                args_ = ([num[c] for c in d], dsh, [num[c] for c in l], Lh, [num[c] for c in r], Rh)
                oracle.contract_labels(*args_)
                reps = int(max(1, min(20, 2e9 // flops)))
                t0 = time.perf_counter()
                for _ in range(reps):
                    oracle.contract_labels(*args_)
                tf.append(flops * reps / (time.perf_counter() - t0) / 1e12)
            per_threads[str(nthr)] = float(np.median(tf))
        gpu_b = float(np.median([x["batch_tflops"] for x in con[str(s_)]]))
        gpu_1 = float(np.median([x["one_block_tflops"] for x in con[str(s_)]]))
        cpu[str(s_)] = {"cpu_tflops_by_threads": per_threads, "gpu_batched_tflops_median": gpu_b, "gpu_one_block_tflops_median": gpu_1,
                        "gpu_batched_over_cpu_best": gpu_b / max(per_threads.values())}
        print(f"config 4 s={s_}: CPU (oracle + OpenBLAS) " + ", ".join(f"{k} thr {v:.3f} TF/s" for k, v in per_threads.items()) +
              f"; GPU batched {gpu_b:.1f} TF/s, one block {gpu_1:.2f} TF/s", flush=True)
    oracle.use_naive_gemm()
    out["config4_vs_cpu"] = {"host_cores": cores, "how": "oracle.contract_labels (permute -> OpenBLAS dgemm -> permute), one block per "
                             "call, median over 4 label patterns", "by_segment_size": cpu}

# ---------------------------------------------------------------------------------------------- permutes
perm = {}
for s in sizes:
    shape = (s, s, s, s)
    n = int(max(1, min(256, (1 << 29) // (8 * s ** 4))))  # ~0.5 GB of blocks per launch
    ins = [api.DeviceBlock(shape).fill(1.0) for _ in range(n)]
    outs = [api.DeviceBlock(shape) for _ in range(n)]
    one, bat = {}, {}
    for p in itertools.permutations(range(4)):
        if p == (0, 1, 2, 3):
            continue
        transp = [1] + [x + 1 for x in p]
        key = "".join(map(str, p))
        one[key] = 16.0 * s ** 4 / time_ms(lambda: api.permute(ins[0], transp, out=outs[0]), reps=7) / 1e6
        bp = api.BatchedPermute(ins, transp, outs)   # pointer arrays marshalled once: time the library, not ctypes
        bat[key] = n * 16.0 * s ** 4 / time_ms(lambda: bp.launch(), reps=3, warm=1) / 1e6
    perm[str(s)] = {"blocks_per_launch": n, "one_block_gbs": one, "batched_gbs": bat}
    vb, v1 = list(bat.values()), list(one.values())
    print(f"permute s={s}: batched({n}) GB/s min/median/max {min(vb):.0f}/{np.median(vb):.0f}/{max(vb):.0f}; one block "
          f"{min(v1):.0f}/{np.median(v1):.0f}/{max(v1):.0f}", flush=True)
    del ins, outs
out["permute_rank4"] = perm

# ---------------------------------------------------------------------------------------------- (T) inner body
tt = {}
for s in ([16, 32] if quick else [16, 32, 50]):
    nij = 16  # ii x jj slices of the rank-6 block  X[a,a1,b,k1,ii,jj]
    X = api.DeviceBlock((s, s, s, s, 4, 4)).fill(0.0)
    t1 = [api.DeviceBlock((s, s, s)).fill_hash(3, i, 1.0) for i in range(nij)]
    t2 = [api.DeviceBlock((s, s, s)).fill_hash(4, i, 1.0) for i in range(nij)]
    # tpppp[a,a1,b,k1] = t1ppp[b1,a1,b] * t2ppp[b1,a,k1]:  labels a=1 a1=2 b=3 k1=4 b1=5
    ptrn, ierr = api.get_contraction_ptrn([1, 2, 3, 4], [5, 2, 3], [5, 1, 4])
    assert ierr == 0
    slices = [X.ptr + 8 * s ** 4 * i for i in range(nij)]
    reps_chain = 8  # 8 accumulations per slice per launch (the b1-segment / k loop of the SIAL body)
    nb = nij * reps_chain
    # chained form: each slice receives the sum of its 8 products and is read-modify-written ONCE
    chained = api.BatchedContraction(ptrn, [(s, s, s)] * nij, [(s, s, s)] * nij, [(s, s, s, s)] * nij,
                                     [t1[i % nij].ptr for i in range(nb)], [t2[(i * 5) % nij].ptr for i in range(nb)],
                                     slices, chain_start=[i * reps_chain for i in range(nij + 1)])
    tmp = [api.DeviceBlock((s, s, s, s)) for _ in range(nij)]
    views = [api.DeviceBlock((s, s, s, s), ptr=p, owned=False) for p in slices]

    def unfused():
        # the reference's order: contract into a temp block, then `X += tmp`, once per product
        for i in range(nb):
            api.contract(ptrn, t1[i % nij], t2[(i * 5) % nij], (s, s, s, s), out=tmp[i % nij])
            views[i // reps_chain].accumulate(tmp[i % nij])

    flops = nb * 2.0 * s ** 5
    ms_c = time_ms(lambda: chained.launch(alpha=1.0, beta=1.0), reps=5)
    ms_u = time_ms(unfused, reps=3, warm=1)
    # algorithmic bytes of the fused form: operands 2*8*s^3 per product + one read-modify-write of each slice
    bytes_c = nb * 16.0 * s ** 3 + nij * 16.0 * s ** 4
    tt[str(s)] = {"products": nb, "chained_fused_ms": ms_c, "unfused_ms": ms_u, "speedup": ms_u / ms_c,
                  "fused_tflops": flops / ms_c / 1e9, "fused_gbs_algorithmic": bytes_c / ms_c / 1e6}
    print(f"(T) s={s}: fused chained {ms_c:.3f} ms ({flops / ms_c / 1e9:.2f} TF/s, {bytes_c / ms_c / 1e6:.0f} GB/s alg.), "
          f"unfused {ms_u:.3f} ms, speed-up {ms_u / ms_c:.1f}x", flush=True)
    del X, t1, t2, tmp, chained
out["triples_inner"] = tt

# ---------------------------------------------------------------------------------------------- EOM-shaped
eom = {}
for (o, v) in ((20, 50), (32, 64)):
    nb = 256
    # R2[k,a,i,b,j] (k = state index, extent 1) contracted with a rank-2 simple-like array: Z[k,a,i] = R2[k,a,i,b,j]*F[j,b]
    R2 = [api.DeviceBlock((1, v, o, v, o)).fill_hash(5, i, 1.0) for i in range(8)]
    F = [api.DeviceBlock((o, v)).fill_hash(6, i, 1.0) for i in range(8)]
    Z = [api.DeviceBlock((1, v, o)) for _ in range(nb)]
    ptrn, ierr = api.get_contraction_ptrn([1, 2, 3], [1, 2, 3, 4, 5], [5, 4])
    assert ierr == 0
    bc = api.BatchedContraction(ptrn, [(1, v, o, v, o)] * nb, [(o, v)] * nb, [(1, v, o)] * nb,
                                [R2[i % 8].ptr for i in range(nb)], [F[i % 8].ptr for i in range(nb)], [z.ptr for z in Z])
    ms = time_ms(lambda: bc.launch(), reps=5)
    bytes_alg = nb * 8.0 * (v * o * v * o + o * v + v * o)
    eom[f"o{o}v{v}"] = {"blocks": nb, "ms": ms, "gbs_algorithmic": bytes_alg / ms / 1e6,
                        "tflops": nb * 2.0 * (v * o) ** 2 / ms / 1e9}
    print(f"EOM-shaped o={o} v={v}: {ms:.3f} ms, {bytes_alg / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    del R2, F, Z, bc
out["eom_shaped"] = eom

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
print("launches", sip.kernel_launches())
