"""ncu target (development): a few batched permute launches of one variant / shape / permutation.
  PERMUTE_BULK=0|1|2 python scripts/ncu_permute.py 32 3210 128"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip
api = sip.api
sip.init(0)
s, perm, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
shape = tuple(int(x) for x in s.split("x")) if "x" in s else (int(s),) * 4
api.set_tuning("permute_bulk", int(os.environ.get("PERMUTE_BULK", "3")))
api.set_tuning("permute_vec", int(os.environ.get("PERMUTE_VEC", "-1")))
ins = [api.DeviceBlock(shape).fill(1.0) for _ in range(n)]
outs = [api.DeviceBlock(shape) for _ in range(n)]
bp = api.BatchedPermute(ins, [1] + [int(c) + 1 for c in perm], outs)
for _ in range(4):
    bp.launch()
api.sync()
