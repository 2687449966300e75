"""BASELINE config 1 ("test/lccd_test.dat LCCD water small basis", the enabled lccd_frozencore_test) and the LCCSD / CCSD jobs ON
THE DEVICE through the reference's own program texts run VERBATIM: tests/golden/rlccd_rhf_program.sialx,
rlccsd_rhf_program.sialx, rccsd_rhf_program.sialx (= src/sialx/qm/cc/*.sialx, scripts/make_cc_program_goldens.py) walked by the
SIAL front-end on libsipgpu -- `DO KITER`, DIIS with its five-index history arrays and scalar-valued DIST_BB contractions,
`energy_denominator_rhf`, the AO ladder over resident AO integrals, the convergence test at the setup's cc_conv -- against
the energy goldens of test/test_qm.cpp.  CPU twin (oracle backend): tests/test_cc_reference_programs_cpu.py."""
import pytest

import device_chain as dc
import lccd_water as lw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


def run(sip, text, case, record):
    inp = lw.inputs(case)
    seg_ext, aoint, fock = dc.hand_over_scf_and_transformation(sip, case, inp)
    l0 = sip.kernel_launches()
    _, _, sc = dc.run_program_on_device(sip, text, case, inp, seg_ext, aoint, fock, record, lw.setup_constants(case),
                                        extra_arrays=dc.static_arrays(sip, seg_ext))
    return sc, sip.kernel_launches() - l0


@pytest.mark.timeout(900, method="thread")
@pytest.mark.parametrize("case,record", [("dat", True), ("fine", False)])
def test_reference_lccd_program_on_the_device(sip, case, record):
    g_corr, g_e, _ = lw.golden(case)
    sc, launches = run(sip, lw.PROGRAM_RLCCD, case, record)
    print(f"\nrlccd_rhf.sialx verbatim on the device ({case}, record={record}): lccd_correlation {sc['lccd_correlation']:.14f} "
          f"(golden {g_corr:.14f}), lccd_energy {sc['lccd_energy']:.14f} (golden {g_e:.14f}), {int(sc['niter'])} iterations, {launches} launches")
    assert abs(sc["lccd_correlation"] - g_corr) < lw.GOLDEN["tolerance"] and abs(sc["lccd_energy"] - g_e) < lw.GOLDEN["tolerance"]
    assert int(sc["niter"]) == 15 and launches > 0


@pytest.mark.timeout(900, method="thread")
def test_config_1_from_ao_integrals_on_the_device(sip):
    """AO integrals + SCF orbitals resident -> tran_rhf_no4v.sialx VERBATIM -> (persistent-array registry: the MO classes never
    leave HBM) -> rlccd_rhf.sialx VERBATIM -> the goldens of lccd_frozencore_test"""
    case = "dat"
    inp = lw.inputs(case)
    seg_ext, aoint, fock = dc.hand_over_scf_and_transformation(sip, case, inp, transformed=False)
    consts = lw.setup_constants(case)
    l0 = sip.kernel_launches()
    dc.run_program_on_device(sip, lw.PROGRAM_TRAN_NO4V, case, inp, seg_ext, aoint, fock, True, consts, extra_arrays=dc.static_arrays(sip, seg_ext))
    _, _, sc = dc.run_program_on_device(sip, lw.PROGRAM_RLCCD, case, inp, seg_ext, aoint, fock, True, consts,
                                        extra_arrays=dc.static_arrays(sip, seg_ext))
    g_corr, g_e, _ = lw.golden(case)
    print(f"\ntran_rhf_no4v.sialx -> rlccd_rhf.sialx verbatim on the device: lccd_correlation {sc['lccd_correlation']:.14f} (golden {g_corr:.14f}), "
          f"{sip.kernel_launches() - l0} launches")
    assert abs(sc["lccd_correlation"] - g_corr) < lw.GOLDEN["tolerance"] and abs(sc["lccd_energy"] - g_e) < lw.GOLDEN["tolerance"]


@pytest.mark.timeout(900, method="thread")
def test_reference_cis_program_on_the_device(sip):
    """tran_rhf_no4v.sialx -> rcis_rhf.sialx VERBATIM on libsipgpu: the CIS roots the reference asserts for this molecule
    (DISABLED_eom_test, test/test_qm.cpp:265-272, 1e-10); `eigen_calc` / `cis_unit_guess` are the host routines they are in the
    reference, everything else -- H-bar, H*B, B*HB matrix elements, residuals, Gram-Schmidt -- runs through the C ABI"""
    case = "eom_dat"
    inp = lw.inputs(case)
    seg_ext, aoint, fock = dc.hand_over_scf_and_transformation(sip, case, inp, transformed=False)
    consts = lw.setup_constants(case)
    dc.run_program_on_device(sip, lw.PROGRAM_TRAN_NO4V, case, inp, seg_ext, aoint, fock, True, consts, extra_arrays=dc.static_arrays(sip, seg_ext))
    w, _, _ = dc.run_program_on_device(sip, lw.PROGRAM_RCIS, case, inp, seg_ext, aoint, fock, True, consts,
                                       extra_arrays=dc.static_arrays(sip, seg_ext))
    roots = [w.tables["sek0"][(k,)] for k in range(1, 5)]
    print(f"\nrcis_rhf.sialx verbatim on the device: CIS roots " + ", ".join(f"{r:.14f}" for r in roots) +
          " (goldens " + ", ".join(f"{r:.14f}" for r in lw.GOLDEN["eom_test"]["cis_sek0"]) + ")")
    for got, want in zip(roots, lw.GOLDEN["eom_test"]["cis_sek0"]):
        assert abs(got - want) < 1e-10, (got, want)


@pytest.mark.timeout(900, method="thread")
def test_reference_lambda_program_on_the_device(sip):
    """the reference's enabled rlambda_test on libsipgpu: tran_rhf_no4v -> rccsd_rhf -> rlambda_rhf verbatim, lambda_pseudo at 1e-10"""
    case = "lam_dat"
    inp = lw.inputs(case)
    seg_ext, aoint, fock = dc.hand_over_scf_and_transformation(sip, case, inp, transformed=False)
    consts = lw.setup_constants(case)
    l0 = sip.kernel_launches()
    for text in (lw.PROGRAM_TRAN_NO4V, lw.PROGRAM_RCCSD, lw.PROGRAM_RLAMBDA):
        _, _, sc = dc.run_program_on_device(sip, text, case, inp, seg_ext, aoint, fock, True, consts, extra_arrays=dc.static_arrays(sip, seg_ext))
    g = lw.GOLDEN["rlambda_test"]["lambda_pseudo"]
    print(f"\ntran -> rccsd -> rlambda verbatim on the device: lambda_pseudo {sc['lambda_pseudo']:.14f} (golden {g:.14f}), {sip.kernel_launches() - l0} launches")
    assert abs(sc["lambda_pseudo"] - g) < 1e-10


@pytest.mark.timeout(900, method="thread")
@pytest.mark.parametrize("case", ["hf_fc_dat", "hf_fc_occ22"])
def test_reference_lambda_ccsdpt_programs_on_the_device(sip, case):
    """the reference's enabled lamccsdpt_test on libsipgpu (hydrogen fluoride / 3-21G, frozen core): tran_rhf_no4v -> rccsd_rhf ->
    rlambda_rhf -> rlamccsdpt_aaa -> rlamccsdpt_aab verbatim; every number the test asserts, at its 1e-10 -- at the setup's
    segmentation and with the active occupied orbitals in two segments (the triples batches cross segment boundaries)"""
    inp = lw.inputs(case)
    g = lw.GOLDEN["lamccsdpt_test"]
    seg_ext, aoint, fock = dc.hand_over_scf_and_transformation(sip, case, inp, transformed=False)
    consts = lw.setup_constants(case)
    l0 = sip.kernel_launches()
    got = {}
    for text in (lw.PROGRAM_TRAN_NO4V, lw.PROGRAM_RCCSD, lw.PROGRAM_RLAMBDA, lw.PROGRAM_RLAMPT_AAA, lw.PROGRAM_RLAMPT_AAB):
        _, _, sc = dc.run_program_on_device(sip, text, case, inp, seg_ext, aoint, fock, True, consts, extra_arrays=dc.static_arrays(sip, seg_ext))
        got.update({k: sc[k] for k in g if k in sc and sc[k] != 0.0})
    print("\nlamccsdpt_test on the device: " + ", ".join(f"{k} {got[k]:.14e} (golden {g[k]:.10e})" for k in g) + f", {sip.kernel_launches() - l0} launches")
    for k in g:
        assert abs(got[k] - g[k]) < 1e-10, (k, got[k], g[k])


@pytest.mark.timeout(900, method="thread")
def test_reference_cis_and_cis_d_programs_on_the_device(sip):
    """the reference's enabled cis_test on libsipgpu (hydrogen fluoride / 3-21G): tran -> rcis_rhf -> rcis_d_rhf verbatim; CIS roots and
    CIS(D) corrections at the test's 1e-10"""
    case = "cis_dat"
    inp = lw.inputs(case)
    g = lw.GOLDEN["cis_test"]
    seg_ext, aoint, fock = dc.hand_over_scf_and_transformation(sip, case, inp, transformed=False)
    consts = lw.setup_constants(case)
    dc.run_program_on_device(sip, lw.PROGRAM_TRAN_NO4V, case, inp, seg_ext, aoint, fock, True, consts, extra_arrays=dc.static_arrays(sip, seg_ext))
    w1, _, _ = dc.run_program_on_device(sip, lw.PROGRAM_RCIS, case, inp, seg_ext, aoint, fock, True, consts,
                                        extra_arrays=dc.static_arrays(sip, seg_ext))
    w2, _, _ = dc.run_program_on_device(sip, lw.PROGRAM_RCIS_D, case, inp, seg_ext, aoint, fock, True, consts,
                                        extra_arrays=dc.static_arrays(sip, seg_ext))
    sek0 = [w1.tables["sek0"][(k,)] for k in (1, 2)]
    ekd = [w2.tables["ekd"][(k,)] for k in (1, 2)]
    print(f"\ncis_test on the device: sek0 {sek0} (goldens {g['sek0']}), ekd {ekd} (goldens {g['ekd']})")
    for got, want in zip(sek0 + ekd, g["sek0"] + g["ekd"]):
        assert abs(got - want) < 1e-10, (sek0, ekd)


@pytest.mark.timeout(900, method="thread")
def test_reference_lccsd_and_ccsd_programs_on_the_device(sip):
    sc, launches = run(sip, lw.PROGRAM_RLCCSD, "all_dat", True)
    g_corr, g_e = lw.golden_lccsd()
    print(f"\nrlccsd_rhf.sialx verbatim on the device: lccsd_correlation {sc['lccsd_correlation']:.14f} (golden {g_corr:.14f}), {launches} launches")
    assert abs(sc["lccsd_correlation"] - g_corr) < 1e-11 and abs(sc["lccsd_energy"] - g_e) < 1e-10
    sc, launches = run(sip, lw.PROGRAM_RCCSD, "eom_dat", True)
    print(f"rccsd_rhf.sialx verbatim on the device: ccsd_energy {sc['ccsd_energy']:.14f} (golden at cc_conv 1e-10 {lw.golden_ccsd()[1]:.14f}), "
          f"{int(sc['niter'])} iterations, {launches} launches")
    assert abs(sc["ccsd_energy"] - lw.golden_ccsd()[1]) < 1e-11 and int(sc["niter"]) == 15
