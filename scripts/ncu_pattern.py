"""ncu target (development): one SIAL contraction pattern as a work-list of nb distinct blocks, a few launches.
  python scripts/ncu_pattern.py 'ab=cade*cbde' oovvo 250 [lowint_scope]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip
api = sip.api
sip.init(0)
pat, kinds, nb = sys.argv[1], sys.argv[2], int(sys.argv[3])
if len(sys.argv) > 4:
    api.set_tuning("lowint_scope", int(sys.argv[4]))
EXT = {"o": 20, "v": 50, "n": 50, "p": 50, "x": 1, "s": 8}
d, rest = pat.split("=")
l, r = rest.split("*")
labs = []
for c in d + l + r:
    if c not in labs:
        labs.append(c)
ext = {c: EXT[k] for c, k in zip(labs, kinds)}
num = {c: i + 1 for i, c in enumerate(labs)}
lsh, rsh, dsh = [ext[c] for c in l], [ext[c] for c in r], [ext[c] for c in d]
ptrn, ierr = api.get_contraction_ptrn([num[c] for c in d], [num[c] for c in l], [num[c] for c in r])
assert ierr == 0
Ls = [api.DeviceBlock(lsh).fill_hash(1, i, 1.0) for i in range(nb)]
Rs = [api.DeviceBlock(rsh).fill_hash(2, i, 1.0) for i in range(nb)]
Ds = [api.DeviceBlock(dsh) for i in range(nb)]
bc = api.BatchedContraction(ptrn, [lsh] * nb, [rsh] * nb, [dsh] * nb, [x.ptr for x in Ls], [x.ptr for x in Rs], [x.ptr for x in Ds])
import torch
stream = torch.cuda.ExternalStream(api.stream_handle())
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    bc.launch()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
byts = 8.0 * nb * (np.prod(lsh) + np.prod(rsh) + np.prod(dsh))
print(pat, kinds, nb, f"{ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s algorithmic")
