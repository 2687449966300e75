"""Where does the host time of a RECORDED CC-program iteration go?  Wraps every libsipgpu entry point the front-end calls
with a wall-clock timer and prints the totals per symbol for one iteration (recorded vs op-at-a-time).
    python scripts/profile_recorded.py [lccsd|ccsd|lccd]"""
import collections
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import aces4_b200  # noqa: E402
import lccd_water as lw  # noqa: E402
from aces4_b200.sial_frontend import DeviceBackend, Program, Walker  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "lccsd"
text, case = {"lccd": (lw.PROGRAM, "fine"), "lccsd": (lw.PROGRAM_LCCSD, "all_fine"), "ccsd": (lw.PROGRAM_CCSD, "all_fine")}[which]
aces4_b200.init()
sip = aces4_b200.api
L = sip.lib()
acc = collections.defaultdict(lambda: [0, 0.0])


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        e = acc[self.name]
        e[0] += 1
        e[1] += time.perf_counter() - t0
        return r


class TimedLib:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        c = self._cache
        if name not in c:
            c[name] = Timed(name, getattr(self._lib, name))
        return c[name]


tl = TimedLib(L)
sip.lib = lambda: tl
for record in (False, True):
    inp = lw.inputs(case)
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    arrays = {}
    for name, kinds in lw.KINDS.items():
        A = sip.DistArray([inp["segs"][k] for k in kinds])
        A.fill_local(0.0)
        for idx, b in inp["arrays"][name].items():
            v = A.block_view(idx)
            sip._check(L.sipgpu_h2d(v.ptr, sip._hp(np.asfortranarray(b)), v.size), "h2d")
        arrays[name] = A
    sip.sync()
    be = DeviceBackend(sip, arrays, record=record)
    be.fock = sip.DeviceBlock.from_numpy(inp["fock"])
    w = Walker(Program(text), be, inp["segs"], index_base=inp["index_base"])
    w.run()
    w.run_proc("iteration")
    sip.sync()
    acc.clear()
    t0 = time.perf_counter()
    w.run_proc("iteration")
    sip.sync()
    dt = time.perf_counter() - t0
    tot = sum(v[1] for v in acc.values())
    print(json.dumps({"program": which, "mode": "recorded" if record else "eager", "iteration_ms": round(dt * 1e3, 1),
                      "in_library_ms": round(tot * 1e3, 1), "stats": sip.wl_stats() if record else None}))
    for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"    {name:36s} calls {n:6d}  total {t * 1e3:9.2f} ms  per call {t / n * 1e6:9.1f} us")
    for A in arrays.values():
        A.destroy()
