"""Energy-level parity report (CPU, oracle backend): the reference's golden energies of test/test_qm.cpp for water /
3-21G (and hydrogen fluoride) next to the values obtained by walking tests/golden/lccd_program.sialx, lccsd_program.sialx
and ccsd_program.sialx block by block with
the SIAL front-end on the CPU oracle.  python scripts/energy_goldens_report.py > profiles/r01_energy_goldens_oracle_backend.txt
(the device rows come from `pytest -m gpu -s tests/test_gpu_z_lccd_water_energy.py`)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lccd_water as lw  # noqa: E402
from aces4_b200.sial_frontend import Program, Walker  # noqa: E402
from oracle import oracle  # noqa: E402
from sial_oracle_backend import OracleBackend  # noqa: E402

oracle.lib()
rows = []
for name, g in ((lw.FROZEN, lw.GOLDEN["scf_energy"]), (lw.ALL, lw.GOLDEN["scf_energy"]),
                ("second_ccsdpt_test.dat", lw.GOLDEN["hf"]["scf_energy"])):
    e = lw.scf(name)[5]
    rows.append((f"scf_energy ({name})", g, e, "numpy input stage (oracle/qm_inputs.py)", "-", "-"))
for program, text, cases in (("lccd", lw.PROGRAM, ("dat", "fine", "all_dat", "all_fine")),
                             ("lccsd", lw.PROGRAM_LCCSD, ("all_dat", "all_fine")),
                             ("ccsd", lw.PROGRAM_CCSD, ("all_dat", "all_fine", "hf_dat", "hf_fine", "hf_fc_dat", "hf_fc_fine",
                                                               "ne_dat", "ne_fine"))):
    for case in cases:
        inp = lw.inputs(case)
        be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
        w = Walker(Program(text), be, inp["segs"], index_base=inp["index_base"])
        t0 = time.time()
        e_mp2, hist = lw.converge(w, be.value, max_iter=150)
        dt = time.time() - t0
        seg = f"occ {inp['segs']['o']} virt {inp['segs']['v']} ao {inp['segs']['ao']}"
        how = f"{program} program, {seg}"
        if program == "lccd":
            g_corr, g_tot, g_mp2 = lw.golden(case)
            if g_corr is not None:
                rows.append(("lccd_correlation (frozen core)", g_corr, hist[-1], how, len(hist), be.calls))
            rows.append((f"lccd_energy ({'frozen core' if g_corr is not None else 'all electron'})", g_tot,
                         hist[-1] + inp["e_scf"], how, len(hist), be.calls))
            if g_mp2 is not None:
                rows.append(("mp2_energy (all electron)", g_mp2, e_mp2 + inp["e_scf"], how, 0, "-"))
        elif program == "ccsd":
            e_tot = hist[-1] + inp["e_scf"]
            if case.startswith("all"):
                g_tight, g_loose = lw.golden_ccsd()
                rows.append(("ccsd_energy water (cc_conv 1e-12)", g_tight, e_tot, how, len(hist), be.calls))
                rows.append(("ccsd_energy water (cc_conv 1e-10)", g_loose, e_tot, how, len(hist), be.calls))
            elif case in ("hf_dat", "hf_fine"):
                gh = lw.GOLDEN["hf"]
                rows.append(("ccsd_correlation HF (cc_conv 1e-10)", gh["ccsd_correlation"], hist[-1], how, len(hist), be.calls))
                rows.append(("ccsd_energy HF (cc_conv 1e-10)", gh["ccsd_energy"], e_tot, how, len(hist), be.calls))
                c0 = be.calls
                sc = Walker(Program(lw.PROGRAM_PT), be, inp["segs"], index_base=inp["index_base"]).run()
                e_t = be.value(sc["et"])
                how_t = how.replace("ccsd program", "ccsd program + restated (T), rank-6 blocks")
                for comp in ("eaaa", "esaaa", "eaab", "esaab"):
                    rows.append((f"{comp} HF", gh[comp], be.value(sc[comp]), how_t, "-", "-"))
                rows.append(("E(T) HF = eaaa+esaaa+eaab+esaab", gh["ccsdpt_energy"] - gh["ccsd_energy"], e_t, how_t, "-", be.calls - c0))
                rows.append(("ccsdpt_energy HF (cc_conv 1e-10)", gh["ccsdpt_energy"], e_tot + e_t, how_t, len(hist), be.calls))
            elif case.startswith("ne"):
                # BASELINE config 2: test/ccsdpt_test.dat (DISABLED_ccsdpt_test, test_qm.cpp:22-52); goldens of a run stopped at
                # the setup's cc_conv 1e-7
                gn = lw.GOLDEN["ne_ccsdpt_test"]
                c0 = be.calls
                sc = Walker(Program(lw.PROGRAM_PT), be, inp["segs"], index_base=inp["index_base"]).run()
                how_t = how.replace("ccsd program", "ccsd program + restated (T), rank-6 blocks")
                for comp in ("eaab", "esaab"):
                    rows.append((f"{comp} Ne ccsdpt_test.dat (cc_conv 1e-7)", gn[comp], be.value(sc[comp]), how_t, len(hist),
                                 be.calls - c0))
            else:
                rows.append(("ccsd_energy HF frozen core (1e-12)", lw.GOLDEN["hf"]["frozen_core_ccsd_energy"], e_tot, how, len(hist), be.calls))
        else:
            g_corr, g_tot = lw.golden_lccsd()
            rows.append(("lccsd_correlation (all electron)", g_corr, hist[-1], how, len(hist), be.calls))
            rows.append(("lccsd_energy (all electron)", g_tot, hist[-1] + inp["e_scf"], how, len(hist), be.calls))
print("reference goldens: test/test_qm.cpp (ASSERT_NEAR 1e-10); north_star tolerance 1e-9 Hartree")
print(f"{'quantity':38s} {'reference golden':>20s} {'reproduced':>20s} {'diff':>9s}  iters  block ops  how")
for q, g, v, how, it, ops in rows:
    print(f"{q:38s} {g:20.14f} {v:20.14f} {v - g:9.1e}  {it!s:>5s}  {ops!s:>9s}  {how}")
