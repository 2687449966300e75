"""Deferred op stream on the REAL coupled-cluster programs, without a GPU: one iteration of tests/golden/lccd_program.sialx,
lccsd_program.sialx, ccsd_program.sialx (+ the integral transformation program and the rank-6 (T) stream) walked on the device backend with the library in dry
mode (fake device addresses; the recorder, the hazard analysis, the fusion passes and the level scheduler run on the
host).  Prints per program: ops recorded -> units scheduled, accumulates fused into their contraction, temporaries elided,
chains (several pairs accumulated into one destination tile walk), levels, and the host time of record + schedule.
python scripts/wl_dry_programs_report.py > profiles/r01_wl_dry_cc_programs.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aces4_b200  # noqa: E402
import lccd_water as lw  # noqa: E402
from aces4_b200.sial_frontend import Program, Walker  # noqa: E402
from test_lccd_program_dry_cpu import DryArray, DryBackend  # noqa: E402

sip = aces4_b200.api
print("one iteration per program, pardo by pardo (every pardo is one recording, as in DeviceBackend(record=True))")
print(f"{'program':8s} {'case':9s} {'recorded':>9s} {'scheduled':>9s} {'fused +=':>9s} {'temps elided':>12s} {'chains':>7s} "
      f"{'chain pairs':>11s} {'levels':>7s} {'pardos':>7s} {'host ms (walker+record+schedule)':>33s}")
for program, text, case in (("tran", lw.PROGRAM_TRAN, "all_fine"), ("lccd", lw.PROGRAM, "fine"), ("lccsd", lw.PROGRAM_LCCSD, "all_fine"),
                            ("ccsd", lw.PROGRAM_CCSD, "all_fine"), ("ccsd", lw.PROGRAM_CCSD, "hf_fine"),
                            ("(T)", lw.PROGRAM_PT, "hf_fine")):
    inp = lw.inputs(case)
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    with sip.recording(dry=True):        # outer dry recording: allocations get fake addresses
        arrays = {n: DryArray(sip, [inp["segs"][k] for k in kinds]) for n, kinds in lw.KINDS.items()}
        be = DryBackend(sip, arrays, record=False)
        be.fock = sip.DeviceBlock(inp["fock"].shape)
        tot = {}
        npardo = [0]
        end_pardo = be.end_pardo

        def flush_and_count():
            sip.wl_flush()
            npardo[0] += 1

        be.end_pardo = flush_and_count
        w = Walker(Program(text), be, inp["segs"], index_base=inp["index_base"])
        if program not in ("(T)", "tran"):
            w.run()
            sip.wl_flush()
        st0 = dict(sip.wl_stats())
        npardo[0] = 0
        t0 = time.perf_counter()
        if program in ("(T)", "tran"):
            w.run()
        else:
            w.run_proc("iteration")
        sip.wl_flush()
        ms = (time.perf_counter() - t0) * 1e3
        st = sip.wl_stats()
        d = {k: st[k] - st0.get(k, 0) for k in st if isinstance(st[k], int)}
    print(f"{program:8s} {case:9s} {d['recorded']:9d} {d['scheduled']:9d} {d['fused_accumulates']:9d} {d['temps_elided']:12d} "
          f"{d['chains']:7d} {d['chain_pairs']:11d} {d['levels']:7d} {npardo[0]:7d} {ms:33.1f}")
