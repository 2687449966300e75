"""Wall time per iteration of the real coupled-cluster programs (tests/golden/*_program.sialx, water / HF in 3-21G) on
libsipgpu: every block operation its own launch (record=False, the naive port of the interpreter) vs the deferred op
stream (record=True).  These molecules are tiny (blocks of 4..625 doubles), so the numbers measure launch and host
overhead, not the kernels -- the regime the work-list exists for.  Prints one JSON line per (program, case, mode).
    python scripts/cc_programs_bench.py            # on a B200
    python scripts/cc_programs_bench.py --fake     # logic check on the CPU against tests/fake_device_api.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import lccd_water as lw  # noqa: E402
from aces4_b200.sial_frontend import DeviceBackend, Program, Walker  # noqa: E402


def main():
    fake = "--fake" in sys.argv
    if fake:
        from fake_device_api import FakeApi
        from oracle import oracle
        oracle.lib()
        sip = FakeApi(oracle)
    else:
        import aces4_b200
        aces4_b200.init()
        sip = aces4_b200.api
    iters = 2 if fake else 5
    for program, text, case in (("lccd", lw.PROGRAM, "fine"), ("lccsd", lw.PROGRAM_LCCSD, "all_fine"),
                                ("ccsd", lw.PROGRAM_CCSD, "all_fine"), ("ccsd", lw.PROGRAM_CCSD, "hf_fine")):
        for record in (False, True):
            inp = lw.inputs(case)
            sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
            arrays = {}
            for name, kinds in lw.KINDS.items():
                A = sip.DistArray([inp["segs"][k] for k in kinds])
                A.fill_local(0.0)
                for idx, b in inp["arrays"][name].items():
                    v = A.block_view(idx)
                    sip._check(sip.lib().sipgpu_h2d(v.ptr, sip._hp(np.asfortranarray(b)), v.size), "h2d")
                arrays[name] = A
            sip.sync()
            be = DeviceBackend(sip, arrays, record=record)
            be.fock = sip.DeviceBlock.from_numpy(inp["fock"])
            w = Walker(Program(text), be, inp["segs"], index_base=inp["index_base"])
            w.run()
            w.run_proc("iteration")          # warm-up
            sip.sync()
            if not fake:
                sip.trace(True)
            l0, t0 = sip.kernel_launches(), time.perf_counter()
            for _ in range(iters):
                e = be.value(w.run_proc("iteration")["ecorrab"])
            sip.sync()
            dt = (time.perf_counter() - t0) / iters
            print(json.dumps({"program": program, "case": case, "segs": {k: inp["segs"][k] for k in ("o", "v", "ao")},
                              "mode": "recorded" if record else "op-at-a-time", "ms_per_iteration": round(dt * 1e3, 2),
                              "launches_per_iteration": (sip.kernel_launches() - l0) // iters, "energy": e,
                              "backend": "fake (CPU)" if fake else "libsipgpu"}), flush=True)
            if not fake:     # per-entry-point table of the timed iterations (the reference's Tracer keeps the same per opcode)
                top = sip.trace_report()[:8]
                sip.trace(False)
                print(json.dumps({"program": program, "mode": "recorded" if record else "op-at-a-time", "trace_top8":
                                  [{"entry": n, "calls_per_iteration": c // iters, "host_ms_per_iteration": round(s * 1e3 / iters, 3)}
                                   for n, c, s in top]}), flush=True)
            for A in arrays.values():
                A.destroy()


if __name__ == "__main__":
    main()
