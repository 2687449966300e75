"""The reference's CCSD program VERBATIM -- tests/golden/rccsd_rhf_program.sialx = src/sialx/qm/cc/rccsd_rhf.sialx, one
documented edit (scripts/make_rccsd_golden.py: the AO integral engine is a `request`) -- walked by the SIAL front-end: the
main program with `DO KITER`, DIIS (MOVET1 / MOVET2 / DIISN: five-index history arrays Daibj[a,i,b,j,kdiis] /
Eaibj[...], the scalar-valued contractions into DIST_BB[jdiis,j1diis], `execute compute_diis BB` = host dgesv), the
convergence test `IF ediff < ecrit ... exit` at the setup's cc_conv and the deferred `set_persistent` hand-over.
Golden: eom_ccsd_water_test's ccsd_energy -75.71251002936883 (test/test_qm.cpp:990-991), which is the value of a run STOPPED
at cc_conv = 1e-10 (the converged energy is -75.71251002928709): reproducing it to 1e-12 means the iteration PATH -- DIIS
extrapolation included -- is the reference's.  Oracle backend (CPU); device twin: tests/test_gpu_z_eom_ccsd.py."""
import numpy as np

import lccd_water as lw
from oracle import qm_inputs as qm
from aces4_b200.sial_frontend import Program, Walker, compute_diis
from sial_oracle_backend import OracleBackend


def run_rccsd(oracle, case):
    """-> (ccsd_energy, iterations, backend calls); leaves the program's persistent arrays in OracleBackend.registry"""
    inp = lw.inputs(case)
    prog = Program(lw.PROGRAM_RCCSD)
    arrays = {n: {} for n in lw.program_array_kinds(prog)}
    arrays["aoint"] = inp["arrays"]["aoint"]
    OracleBackend.registry.clear()
    OracleBackend.registry.update({lab: inp["arrays"][lab.lower()] for lab in lw.PERSISTED})      # the transformation program's
    OracleBackend.registry.update(scf_energy=inp["e_scf"], ca=inp["arrays"]["ca"],
                                  fock_a=qm.split_blocks(inp["fock"], [inp["segs"]["p"], inp["segs"]["p"]]))
    be = OracleBackend(oracle, arrays, fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(prog, be, inp["segs"], index_base=inp["index_base"], constants=lw.eom_constants())
    sc = w.run()
    return be.value(sc["ccsd_energy"]), int(be.value(sc["niter"])), be.calls


def test_reference_ccsd_program_with_diis_stops_where_the_reference_stops(oracle):
    """measured: -75.71251002936886 after 15 iterations (golden -75.71251002936883: 3e-14)"""
    e, niter, calls = run_rccsd(oracle, "eom_dat")
    assert abs(e - lw.golden_ccsd()[1]) < 1e-12, e
    assert niter == 15
    reg = OracleBackend.registry
    assert {"t1a_old", "T2old_aa", "T2old_ab", "VSpipi", "Vaaii", "Viaai", "Vaaai", "Vpiqj", "ca", "fock_a", "ccsd_energy"} <= set(reg)
    assert abs(reg["ccsd_energy"] - e) == 0.0 and reg["has_singles"] == 1.0


def test_reference_ccsd_program_at_a_finer_segmentation(oracle):
    """occupied 2 + 3, virtual 3 + 5, AO 4 + 7 + 2: the same iteration path block by block"""
    e, niter, calls = run_rccsd(oracle, "eom_fine")
    assert abs(e - lw.golden_ccsd()[1]) < 1e-12 and niter == 15


def test_compute_diis_follows_form_R():
    """form_R.F: upper triangle symmetrised, trailing all-zero rows dropped, bordered system solved; the coefficients sum to 1"""
    rng = np.random.default_rng(4)
    n, m = 6, 3
    E = rng.uniform(-1, 1, (m, 20))
    B = np.zeros((n, n))
    B[:m, :m] = np.triu(E @ E.T)          # only the upper triangle is read
    c = compute_diis(B.tolist())
    assert abs(sum(c) - 1.0) < 1e-12 and np.all(c[m:] == 0.0)
    full = E @ E.T
    M = np.block([[full, -np.ones((m, 1))], [-np.ones((1, m)), np.zeros((1, 1))]])
    ref = np.linalg.solve(M, np.r_[np.zeros(m), -1.0])[:m]
    assert np.allclose(c[:m], ref, atol=1e-13)
