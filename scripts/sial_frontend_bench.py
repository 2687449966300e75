"""LCCD doubles SIAL fragment (tests/golden/lccd_doubles.sialx) through the SIAL front-end on one GPU: the per-block
call stream executed op-at-a-time vs recorded and batched by the deferred op stream (worklist.cu).  Prints one JSON
line: wall time (host + device, synchronised), kernel launches, algorithmic TFLOP/s and the scheduler statistics."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aces4_b200 as sip  # noqa: E402
from aces4_b200.sial_frontend import DeviceBackend, Program, Walker  # noqa: E402

api = sip.api
KINDS = {"vpiqj": "vovo", "voooo": "oooo", "viaai": "ovvo", "vaaii": "vvoo", "t2old_ab": "vovo", "t2new_ab": "vovo"}
TAGS = {"vpiqj": 2, "voooo": 3, "viaai": 5, "vaaii": 6, "t2old_ab": 1, "t2new_ab": 9}


def main():
    o_segs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,20").split(",")]
    v_segs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "50,50,50,50").split(",")]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    segs = {"o": o_segs, "v": v_segs}
    sip.init(0)
    text = open(os.path.join(ROOT, "tests", "golden", "lccd_doubles.sialx")).read()
    prog = Program(text)
    o, v = sum(o_segs), sum(v_segs)
    flops = 2.0 * o ** 4 * v ** 2 + 3 * 2.0 * o ** 3 * v ** 3
    out = {"workload": f"lccd_doubles.sialx o={o} ({len(o_segs)} segs) v={v} ({len(v_segs)} segs)", "flops": flops}
    energies = {}
    for mode, record in (("op_at_a_time", False), ("recorded", True)):
        arrays = {}
        for name, kind in KINDS.items():
            A = api.DistArray([segs[k] for k in kind])
            nseg = [len(segs[k]) for k in kind]
            for idx in np.ndindex(*nseg):
                idx1 = tuple(i + 1 for i in idx)
                b = A.block_view(idx1)
                if name == "t2new_ab":
                    b.fill(0.0)
                else:
                    b.fill_hash(0xACE54, (TAGS[name] << 40) | A.block_number(idx1), 0.1)
            arrays[name] = A
        api.sync()
        times, launches, stats = [], [], None
        for r in range(reps + 1):
            arrays["t2new_ab"].fill_local(0.0)
            api.sync()
            be = DeviceBackend(api, arrays, record=record)
            l0 = sip.kernel_launches()
            t0 = time.perf_counter()
            scal = Walker(prog, be, segs).run()
            e = be.value(scal["ecorrab"])
            api.sync()
            t1 = time.perf_counter()
            if r > 0:
                times.append(t1 - t0)
                launches.append(sip.kernel_launches() - l0)
            stats = be.stats
        energies[mode] = e
        sec = min(times)
        out[mode] = {"seconds": sec, "tflops": flops / sec / 1e12, "kernel_launches": launches[-1], "energy": e}
        if record:
            keys = stats[0].keys()
            out[mode]["worklist"] = {k: int(sum(s[k] for s in stats)) for k in keys}
        for A in arrays.values():
            A.destroy()
    out["energy_rel_diff"] = abs(energies["recorded"] - energies["op_at_a_time"]) / abs(energies["op_at_a_time"])
    out["speedup"] = out["op_at_a_time"]["seconds"] / out["recorded"]["seconds"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
