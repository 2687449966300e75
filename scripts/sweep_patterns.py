"""Every distinct contraction label pattern of the reference's SIAL programs (tests/golden/sial_contraction_patterns.txt,
170 patterns: 68x(4,4,2), 41x(4,4,4), 18x(2,4,4), 16x(2,4,2), 16x(2,2,2), ...) at CCSD block extents (occupied 20,
virtual/ao 50, simple index 1), timed as a work-list of nb blocks per launch with DISTINCT operand blocks (so that the
bandwidth-bound shapes really stream from HBM).  Per pattern: TFLOP/s, algorithmic GB/s, and the fraction of the
roofline time  max(flops / DMMA peak, algorithmic bytes / copy bandwidth).  Writes gpurun_out/sweep_patterns.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aces4_b200 as sip  # noqa: E402
from conftest import sial_patterns  # noqa: E402

api = sip.api
sip.init(0)
torch.cuda.set_device(0)
stream = torch.cuda.ExternalStream(api.stream_handle())
EXT = {"o": 20, "v": 50, "n": 50, "p": 50, "x": 1, "s": 8}


def time_ms(fn, reps=3, warm=1):
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


peak = max(api.dmma_peak_probe(40000) for _ in range(2))
bw = api.copy_bw_probe(1 << 30, 10)
rows = []
for d, l, r, kinds, where in sial_patterns():
    if not d:
        continue  # scalar-valued contractions are dot products (elementwise.cu), not this kernel
    if os.environ.get("SWEEP_RANKS") and f"{len(d)}{len(l)}{len(r)}" not in os.environ["SWEEP_RANKS"].split(","):
        continue
    labs = []
    for c in d + l + r:
        if c not in labs:
            labs.append(c)
    num = {c: i + 1 for i, c in enumerate(labs)}
    ext = {c: EXT[kinds[c]] for c in labs}
    lsh, rsh, dsh = [ext[c] for c in l], [ext[c] for c in r], [ext[c] for c in d]
    K = float(np.prod([ext[c] for c in l if c in r]))
    flops = 2.0 * np.prod(dsh) * K
    byts = 8.0 * (np.prod(lsh) + np.prod(rsh) + np.prod(dsh))
    ptrn, ierr = api.get_contraction_ptrn([num[c] for c in d], [num[c] for c in l], [num[c] for c in r])
    if ierr != 0:
        continue
    nb = int(max(8, min(int(os.environ.get("SWEEP_NB_CAP", "1024")), max(np.ceil(10e9 / flops), np.ceil(1.5e9 / byts)))))
    def pool(shape, tag):
        n = int(max(1, min(nb, 1.5e9 // (8 * np.prod(shape)))))
        return [api.DeviceBlock(shape).fill_hash(tag, i, 1.0) for i in range(n)]
    Ls, Rs, Ds = pool(lsh, 1), pool(rsh, 2), pool(dsh, 3)
    bc = api.BatchedContraction(ptrn, [lsh] * nb, [rsh] * nb, [dsh] * nb, [Ls[i % len(Ls)].ptr for i in range(nb)],
                                [Rs[i % len(Rs)].ptr for i in range(nb)], [Ds[i % len(Ds)].ptr for i in range(nb)])
    ms = time_ms(lambda: bc.launch())
    t_roof = max(nb * flops / (peak * 1e12), nb * byts / (bw * 1e9)) * 1e3
    rows.append({"pattern": f"{d}={l}*{r}", "kinds": "".join(kinds[c] for c in labs), "ranks": f"{len(d)}{len(l)}{len(r)}",
                 "where": where, "M*N": float(np.prod(dsh)), "K": K, "blocks": nb, "ms": ms,
                 "tflops": nb * flops / ms / 1e9, "gbs_algorithmic": nb * byts / ms / 1e6,
                 "bound": "tensor" if nb * flops / (peak * 1e12) > nb * byts / (bw * 1e9) else "hbm",
                 "roofline_frac": t_roof / ms})
    del Ls, Rs, Ds, bc
out = {"dmma_peak_tflops": peak, "copy_gbs": bw, "extents": EXT, "patterns": rows}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("SWEEP_OUT", "sweep_patterns.json")), "w"), indent=1)
for cls in sorted(set(x["ranks"] for x in rows)):
    sel = [x for x in rows if x["ranks"] == cls]
    f = [x["roofline_frac"] for x in sel]
    nt = sum(1 for x in sel if x["bound"] == "tensor")
    print(f"ranks {cls}: {len(sel):3d} patterns ({nt} tensor-bound), roofline fraction min/median/max "
          f"{min(f):.2f}/{np.median(f):.2f}/{max(f):.2f}; TF/s median {np.median([x['tflops'] for x in sel]):.1f}, "
          f"GB/s median {np.median([x['gbs_algorithmic'] for x in sel]):.0f}")
worst = sorted(rows, key=lambda x: x["roofline_frac"])[:12]
for x in worst:
    print(f"  worst: {x['pattern']:24s} {x['kinds']:8s} {x['where']:34s} K={x['K']:.0f} MN={x['M*N']:.0f} {x['bound']:6s} "
          f"{x['roofline_frac']:.2f}  {x['tflops']:.2f} TF/s {x['gbs_algorithmic']:.0f} GB/s")
