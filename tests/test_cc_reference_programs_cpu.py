"""The reference's coupled-cluster programs VERBATIM -- tests/golden/rlccd_rhf_program.sialx, rlccsd_rhf_program.sialx,
rccsd_rhf_program.sialx = src/sialx/qm/cc/rlccd_rhf.sialx, rlccsd_rhf.sialx, rccsd_rhf.sialx with the documented edits of
scripts/make_cc_program_goldens.py (the AO integral engine is a `request`; LCCD's response-dipole call left out) -- walked by the
SIAL front-end: the main program with `DO KITER`, DIIS (MOVET1 / MOVET2 / DIISN: five-index history arrays
Daibj[a,i,b,j,kdiis] / Eaibj[...], the scalar-valued contractions into DIST_BB[jdiis,j1diis], `execute compute_diis BB` = host
dgesv), the convergence test `IF ediff < ecrit ... exit` at the setup's cc_conv and the deferred `set_persistent` hand-over.
Goldens (test/test_qm.cpp): BASELINE config 1 -- lccd_frozencore_test's lccd_correlation / lccd_energy (:459-462, the enabled
test; frozen core: `ca` / `fock_a` over ALL orbital segments, the active ranges start at segment 2), the all-electron LCCD
energy (:677-678), lccsd_test (:526-529), and eom_ccsd_water_test's ccsd_energy -75.71251002936883 (:990-991), which is the
value of a run STOPPED at cc_conv = 1e-10 (the converged energy is -75.71251002928709): reproducing it to 1e-12 means the
iteration PATH -- DIIS extrapolation included -- is the reference's.  Oracle backend (CPU); device twins:
tests/test_gpu_z_eom_ccsd.py, tests/test_gpu_z_lccd_water_energy.py."""
import os

import numpy as np
import pytest

import lccd_water as lw
from oracle import qm_inputs as qm
from aces4_b200.sial_frontend import Program, Walker, compute_diis
from aces4_b200.sial_frontend import Walker as W
from sial_oracle_backend import OracleBackend

PROGRAMS = {"rccsd_rhf": lw.PROGRAM_RCCSD, "rlccd_rhf": lw.PROGRAM_RLCCD, "rlccsd_rhf": lw.PROGRAM_RLCCSD,
            "tran_rhf_no4v": lw.PROGRAM_TRAN_NO4V, "rcis_rhf": lw.PROGRAM_RCIS, "rlambda_rhf": lw.PROGRAM_RLAMBDA,
            "rlamccsdpt_aaa": lw.PROGRAM_RLAMPT_AAA, "rlamccsdpt_aab": lw.PROGRAM_RLAMPT_AAB, "rcis_d_rhf": lw.PROGRAM_RCIS_D}


def run_cc_program(oracle, name, case, chained=False):
    """-> (walker scalars as floats, backend calls); leaves the program's persistent arrays in OracleBackend.registry.
    chained: the registry already holds what the program before it handed over (else: the SCF results and the dense numpy
    transformation of oracle/qm_inputs.py are put there)"""
    inp = lw.inputs(case)
    prog = Program(PROGRAMS[name])
    arrays = {n: {} for n in lw.program_array_kinds(prog)}
    arrays["aoint"] = inp["arrays"]["aoint"]
    if not chained:
        OracleBackend.registry.clear()
        if name != "tran_rhf_no4v":
            OracleBackend.registry.update({lab: inp["arrays"][lab.lower()] for lab in lw.PERSISTED})      # the transformation program's
        OracleBackend.registry.update(scf_energy=inp["e_scf"], **lw.all_orbital_statics(case, inp))
    be = OracleBackend(oracle, arrays, fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    host_data, scf_dipole = lw.dipole_data(lw.CASES[case][0])
    Walker.host_registry.setdefault("scf_dipole", scf_dipole)
    w = Walker(prog, be, lw.segs_with_all_orbitals(inp), index_base=inp["index_base"], constants=lw.setup_constants(case),
               host_data=host_data)
    sc = w.run()
    out = {k: be.value(v) for k, v in sc.items()}
    out["tables"] = w.tables
    return out, be.calls


def run_rccsd(oracle, case):
    sc, calls = run_cc_program(oracle, "rccsd_rhf", case)
    return sc["ccsd_energy"], int(sc["niter"]), calls


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


@pytest.mark.parametrize("case", ["dat", "fine"])
def test_reference_lccd_program_reproduces_lccd_frozencore_test(oracle, case):
    """BASELINE config 1 through the reference's own program text (frozen core: moa [1 | 4 | 8]; `fine`: [1 | 2 2 | 3 5]).
    measured: lccd_correlation -0.12610179885837 (golden -0.12610179886435: 6.0e-12), lccd_energy -75.71042854159870 (6.1e-12),
    15 iterations (the setup's cc_conv is 1e-12)"""
    g_corr, g_e, _ = lw.golden(case)
    sc, calls = run_cc_program(oracle, "rlccd_rhf", case)
    assert abs(sc["lccd_correlation"] - g_corr) < lw.GOLDEN["tolerance"], sc["lccd_correlation"]
    assert abs(sc["lccd_energy"] - g_e) < lw.GOLDEN["tolerance"], sc["lccd_energy"]
    assert abs(sc["lccd_correlation"] - g_corr) < 2e-11 and int(sc["niter"]) == 15


@pytest.mark.parametrize("case", ["dat", "fine"])
def test_config_1_from_ao_integrals_through_the_reference_texts(oracle, case):
    """BASELINE config 1 end to end on the hot path: AO integrals + SCF orbitals -> tran_rhf_no4v.sialx VERBATIM (four quarter
    transformations, the six MO classes handed over with set_persistent) -> rlccd_rhf.sialx VERBATIM -> the goldens of
    lccd_frozencore_test.  The transformed classes equal the dense numpy transformation to 3e-15."""
    inp = lw.inputs(case)
    run_cc_program(oracle, "tran_rhf_no4v", case)
    reg = OracleBackend.registry
    assert {"VSpipi", "Vaaii", "Viaai", "Vaaai", "VSaaai", "Vpiqj", "ca"} <= set(reg)
    for lab in lw.PERSISTED:
        ref = inp["arrays"][lab.lower()]
        assert set(reg[lab]) == set(ref)
        assert max(float(np.max(np.abs(reg[lab][k] - ref[k]))) for k in ref) < 1e-13, lab
    sc, _ = run_cc_program(oracle, "rlccd_rhf", case, chained=True)
    g_corr, g_e, _ = lw.golden(case)
    assert abs(sc["lccd_correlation"] - g_corr) < 2e-11 and abs(sc["lccd_energy"] - g_e) < 2e-11 and int(sc["niter"]) == 15


def test_reference_cis_program_reproduces_the_cis_roots_of_eom_test(oracle):
    """rcis_rhf.sialx VERBATIM (scripts/make_eom_golden.py: transition-dipole part left out): H-bar of CIS, `cis_unit_guess`,
    the subspace-collapse Davidson with `eigen_calc` (host dsyev), `invert_diagonal`, `return_diagonal_elements`, contiguous
    local arrays addressed block-wise through int variables -- against the two CIS roots the reference asserts for this molecule
    (DISABLED_eom_test, test/test_qm.cpp:265-272, 1e-10).  measured: 4.3e-13, 2.7e-13; roots 3, 4 equal the dense
    diagonalisation of oracle/qm_inputs.py::cis_singlets to 1e-11"""
    case = "eom_dat"
    inp = lw.inputs(case)
    run_cc_program(oracle, "tran_rhf_no4v", case)
    sc, calls = run_cc_program(oracle, "rcis_rhf", case, chained=True)
    roots = [sc["tables"]["sek0"][(k,)] for k in range(1, 5)]
    for got, want in zip(roots, lw.GOLDEN["eom_test"]["cis_sek0"]):
        assert abs(got - want) < 1e-10, (got, want)
    dense = {n: qm.join_blocks(inp["arrays"][n], [inp["segs"][k] for k in lw.KINDS[n]]) for n in ("vpiqj", "vaaii")}
    e_dense, _ = lw.cis_guess(inp, dense)
    assert max(abs(a - b) for a, b in zip(roots, e_dense)) < 1e-10
    assert {"C1_a", "B1_a", "Vaaii", "Viaai", "ca", "fock_a"} <= set(OracleBackend.registry)
    assert abs(W.host_registry["CIS_E"][(1,)] - roots[0]) == 0.0


@pytest.mark.parametrize("case", ["lam_dat"] + (["lam_fine"] if os.environ.get("SIPGPU_SLOW_TESTS") else []))
def test_reference_lambda_program_reproduces_rlambda_test(oracle, case):
    """The reference's ENABLED rlambda_test (test/test_qm.cpp:307-341; hydrogen fluoride / 3-21G): scf -> tran_rhf_no4v ->
    rccsd_rhf -> rlambda_rhf, every program the reference's text (scripts/make_cc_program_goldens.py; `compute_dipole_integrals` served
    from a resident table): 26 intermediates (F1ae, F1mi, Gae, Gmi, W1minj, W2mebj, W1imen, W1eafm ...), the lambda ladder over
    the AO integrals, DIIS, `lambda_pseudo` asserted at 1e-10.  measured: -0.12592115116562566 vs -0.12592115116563 (4e-15),
    17 iterations"""
    run_cc_program(oracle, "tran_rhf_no4v", case)
    sc, _ = run_cc_program(oracle, "rccsd_rhf", case, chained=True)
    assert abs(sc["ccsd_energy"] - lw.GOLDEN["hf"]["ccsd_energy"]) < 1e-10      # second_ccsdpt_test's CCSD energy: the same molecule
    W.host_registry.clear()
    sc, calls = run_cc_program(oracle, "rlambda_rhf", case, chained=True)
    g = lw.GOLDEN["rlambda_test"]
    assert abs(sc["lambda_pseudo"] - g["lambda_pseudo"]) < 1e-12, sc["lambda_pseudo"]
    # form_G1: the CCSD response density (DABA, DIJA, DIAA, DAIA: block contractions of lambda and T amplitudes), back-transformed
    # and traced with the dipole integrals.  The reference asserts x = y = 0 at 1e-6 and carries z in its `expected` array
    dip = sc["tables"]["dipole"]
    assert abs(dip[(1,)]) < 1e-10 and abs(dip[(2,)]) < 1e-10
    assert abs(dip[(3,)] - g["ccsd_dipole_z"]) < 1e-10, dip           # measured: 8e-15
    assert abs(lw.dipole_data("rlambda_test.dat")[1][(3,)] - g["scf_dipole_z"]) < 1e-9
    assert abs(W.host_registry["ccsd_dipole"][(3,)] - dip[(3,)]) == 0.0
    assert {"l1a_old", "L2old_aa", "L2old_ab", "t1a_old", "T2old_ab"} <= set(OracleBackend.registry)


@pytest.mark.parametrize("case", ["hf_fc_dat", "hf_fc_virt_fine", "hf_fc_occ22"])
def test_reference_lambda_ccsdpt_programs_reproduce_lamccsdpt_test(oracle, case):
    """The reference's ENABLED lamccsdpt_test (test/test_qm.cpp:798-869; hydrogen fluoride / 3-21G, frozen core): scf ->
    tran_rhf_no4v -> rccsd_rhf -> rlambda_rhf -> rlamccsdpt_aaa -> rlamccsdpt_aab, every program the reference's text with the one
    edit of scripts/make_cc_program_goldens.py (their own TRAN_UHF transformation procedures included): Lambda-CCSD(T) -- the
    stripi / one-segment-contraction / rank-6 accumulate inner loops of the (T) programs with lambda amplitudes on the left.
    Every number the test asserts (1e-10).  measured: ccsd_energy 7e-13, eaaa 5e-16, esaaa 1e-15, eaab 2e-15, esaab 1e-16,
    ccsdpt_energy 1e-12.  hf_fc_virt_fine: the virtual space cut into 2 + 4 (AO 3 + 6 + 2), the occupied space in ONE segment like in
    every setup the reference ships.  hf_fc_occ22: the four active occupied orbitals in TWO segments of two (eaaa 5e-16, esaaa 1.5e-15:
    set_ijk_aaa pieces across segment boundaries are fine).  The 1e-7 deviation seen earlier at a 3 + 1 cut is the one-orbital
    segment: energy_denominator_rhf.F takes any extent-1 dimension of a rank-6 block for a simple index."""
    g = lw.GOLDEN["lamccsdpt_test"]
    W.host_registry.clear()
    run_cc_program(oracle, "tran_rhf_no4v", case)
    sc, _ = run_cc_program(oracle, "rccsd_rhf", case, chained=True)
    assert abs(sc["ccsd_energy"] - g["ccsd_energy"]) < 1e-10
    run_cc_program(oracle, "rlambda_rhf", case, chained=True)
    got, _ = run_cc_program(oracle, "rlamccsdpt_aaa", case, chained=True)
    sc, calls = run_cc_program(oracle, "rlamccsdpt_aab", case, chained=True)
    got = {"eaaa": got["eaaa"], "esaaa": got["esaaa"], "eaab": sc["eaab"], "esaab": sc["esaab"], "ccsdpt_energy": sc["ccsdpt_energy"]}
    for k, v in got.items():
        assert abs(v - g[k]) < 1e-10, (k, v, g[k])
        assert abs(v - g[k]) < (1e-11 if k == "ccsdpt_energy" else 1e-13), (k, v, g[k])


def test_reference_cis_and_cis_d_programs_reproduce_cis_test(oracle):
    """The reference's ENABLED cis_test (test/test_qm.cpp:153-202; hydrogen fluoride / 3-21G, the degenerate 1-Pi pair): scf ->
    tran -> rcis_rhf -> rcis_d_rhf verbatim: the CIS roots `sek0` and the CIS(D) corrections `ekd` (the doubles correction with the
    shifted denominator `energy_ty_denominator_rhf`), all at 1e-10.  measured: sek0 2.7e-12 / 9e-13, ekd 1.3e-13 / 1.6e-13.
    (The reference's job transforms with tran_rhf_no3v; tran_rhf_no4v produces the same classes the two programs restore.)"""
    case = "cis_dat"
    g = lw.GOLDEN["cis_test"]
    W.host_registry.clear()
    run_cc_program(oracle, "tran_rhf_no4v", case)
    reg = OracleBackend.registry
    kept = {lab: reg[lab] for lab in ("Vpiqj", "VSpipi")}        # (rcis_rhf does not touch them; rcis_d_rhf restores them)
    sc, _ = run_cc_program(oracle, "rcis_rhf", case, chained=True)
    for k in (1, 2):
        assert abs(sc["tables"]["sek0"][(k,)] - g["sek0"][k - 1]) < 1e-10, sc["tables"]["sek0"]
    reg.update({lab: a for lab, a in kept.items() if lab not in reg})
    sc, _ = run_cc_program(oracle, "rcis_d_rhf", case, chained=True)
    for k in (1, 2):
        assert abs(sc["tables"]["ekd"][(k,)] - g["ekd"][k - 1]) < 1e-10, sc["tables"]["ekd"]
        assert abs(sc["tables"]["ekd"][(k,)] - g["ekd"][k - 1]) < 1e-12


def test_reference_lccd_and_lccsd_programs_all_electron(oracle):
    """eom_lccd_test / lccsd_test (test/test_qm.cpp:677-678, 526-529).  measured: lccd_energy 2.9e-12, lccsd_correlation
    -0.12865706498546847 vs the golden -0.12865706498547 (1.5e-15), lccsd_energy 1.3e-13"""
    sc, _ = run_cc_program(oracle, "rlccd_rhf", "all_dat")
    assert abs(sc["lccd_energy"] - lw.golden("all_dat")[1]) < 2e-11
    sc, _ = run_cc_program(oracle, "rlccsd_rhf", "all_dat")
    g_corr, g_e = lw.golden_lccsd()
    assert abs(sc["lccsd_correlation"] - g_corr) < 1e-12 and abs(sc["lccsd_energy"] - g_e) < 1e-11


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
