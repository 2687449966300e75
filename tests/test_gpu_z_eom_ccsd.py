"""BASELINE config 3 ("test/eom_ccsd_water_test.dat EOM-CCSD water, Davidson sigma-vector contractions") ON THE DEVICE
through the reference's own right-hand EOM-CCSD program: tests/golden/eom_ccsd_right_program.sialx (= src/sialx/qm/eom/
eom_ccsd_rhf_right.sialx + eom_rhf_hbar.sialx, see scripts/make_eom_golden.py) walked by the SIAL front-end on libsipgpu
after the reference's CCSD program run VERBATIM (tests/golden/rccsd_rhf_program.sialx: DIIS with its five-index history arrays
and scalar-valued DIST_BB contractions, stopped at the setup's cc_conv with the golden's ccsd_energy), chained through the
library's persistent-array registry: H-bar, its diagonal, and per Davidson step the sigma-vector contractions on rank-5
blocks with a leading simple index, the subspace matrix elements (scalar-valued contractions), the preconditioned residual
(`invert_diagonal`, `invert_diagonal_asym`), `anti_symm_o/v`, `return_diagonal_elements` and all put / get / `prepare *=`
traffic of the subspace arrays are C-ABI calls; the <= 60 x 60 `gen_eigen_calc` (dgeev) stays on the host as in the
reference.  Goldens: the four roots of eom_ccsd_water_test (test/test_qm.cpp:1005-1013, 1e-8).  CPU twin (oracle backend):
tests/test_eom_ccsd_cpu.py."""
import numpy as np
import pytest

import lccd_water as lw
from oracle import qm_inputs as qm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


def upload(sip, A, blocks):
    for idx, b in blocks.items():
        view = A.block_view(idx)
        sip._check(sip.lib().sipgpu_h2d(view.ptr, sip._hp(np.asfortranarray(b)), view.size), "h2d")


def run_eom_on_device(sip, case, record):
    """the reference's chain on libsipgpu: rccsd_rhf.sialx verbatim (DIIS, stopped at the setup's cc_conv) -> persistent arrays
    (the library's label registry: the slabs never leave HBM) -> the EOM program"""
    from aces4_b200.sial_frontend import DeviceBackend, Program, Walker

    inp = lw.inputs(case)
    consts = lw.eom_constants()
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    seg_ext = dict(inp["segs"])
    seg_ext["p"] = list(inp["segs"]["o"]) + list(inp["segs"]["v"])
    fock = sip.DeviceBlock.from_numpy(inp["fock"])

    def resident(name, kinds, blocks):
        A = sip.DistArray([seg_ext[k] for k in kinds])
        A.fill_local(0.0)
        upload(sip, A, blocks)
        return A

    # what the SCF / transformation programs hand over
    given = {lab: resident(lab, lw.KINDS[lab.lower()], inp["arrays"][lab.lower()]) for lab in lw.PERSISTED}
    given["ca"] = resident("ca", lw.KINDS["ca"], inp["arrays"]["ca"])
    given["fock_a"] = resident("fock_a", ("p", "p"), qm.split_blocks(inp["fock"], [seg_ext["p"], seg_ext["p"]]))
    aoint = resident("aoint", lw.KINDS["aoint"], inp["arrays"]["aoint"])
    sip.sync()
    for label, A in given.items():
        A.persist(label)
    sip.persist_scalar("scf_energy", inp["e_scf"])

    # ---- CCSD: the reference's program, verbatim ----
    prog = Program(lw.PROGRAM_RCCSD)
    arr = lw.device_program_arrays(sip, prog, lw.PROGRAM_RCCSD, consts, inp["segs"], skip=("aoint",))
    arr["aoint"] = aoint
    for name in ("ca", "fock_a"):
        arr[name] = sip.DistArray([seg_ext[k] for k in (lw.KINDS["ca"] if name == "ca" else ("p", "p"))])
    be = DeviceBackend(sip, arr, record=record)
    be.fock, be.seg_ranges = fock, inp["moa_seg_ranges"]
    l0 = sip.kernel_launches()
    sc = Walker(prog, be, inp["segs"], index_base=inp["index_base"], constants=consts).run()
    e_ccsd, niter = be.value(sc["ccsd_energy"]), int(be.value(sc["niter"]))

    # ---- what rlambda / rcis leave behind: VSaaai and the CIS vectors ----
    frag = {"vaaai": sip.DistArray([seg_ext[k] for k in lw.KINDS["vaaai"]]), "vsaaai": sip.DistArray([seg_ext[k] for k in lw.KINDS["vsaaai"]])}
    frag["vaaai"].restore("Vaaai")
    frag["vsaaai"].fill_local(0.0)
    bf = DeviceBackend(sip, frag, record=record)
    Walker(Program(lw.VSAAAI_FRAGMENT), bf, inp["segs"], index_base=inp["index_base"]).run()
    sip.sync()
    frag["vaaai"].persist("Vaaai")
    frag["vsaaai"].persist("VSaaai")
    dense = {}
    for n, lab in (("vpiqj", "Vpiqj"), ("vaaii", "Vaaii")):
        A = sip.DistArray([seg_ext[k] for k in lw.KINDS[n]])
        A.restore(lab)
        dense[n] = qm.join_blocks({idx: A.block_view(idx).to_numpy() for idx in inp["arrays"][n]}, [inp["segs"][k] for k in lw.KINDS[n]])
        A.persist(lab)
    e_cis, c1 = lw.cis_guess(inp, dense)

    # ---- EOM-CCSD: the reference's right-hand program ----
    prog2 = Program(lw.PROGRAM_EOM)
    parr = lw.device_program_arrays(sip, prog2, lw.PROGRAM_EOM, consts, inp["segs"], skip=("aoint",))
    upload(sip, parr["c1_a"], c1)
    sip.sync()
    parr["c1_a"].persist("C1_a")
    parr["aoint"] = aoint
    for name in ("ca", "fock_a"):
        parr[name] = sip.DistArray([seg_ext[k] for k in (lw.KINDS["ca"] if name == "ca" else ("p", "p"))])
        parr[name].restore(name)
    be2 = DeviceBackend(sip, parr, record=record)
    be2.fock, be2.seg_ranges = fock, inp["moa_seg_ranges"]
    w2 = Walker(prog2, be2, inp["segs"], index_base=inp["index_base"], constants=consts)
    w2.run()
    roots = [w2.tables["sek0"][(k,)] for k in range(1, len(e_cis) + 1)]
    return roots, e_cis, e_ccsd, niter, sip.kernel_launches() - l0


@pytest.mark.timeout(1500, method="thread")
@pytest.mark.parametrize("case,record", [("eom_dat", True), ("eom_fine", False)])
def test_reference_eom_program_on_the_device(sip, case, record):
    g = lw.GOLDEN["eom_ccsd_water_test"]
    roots, e_cis, e_ccsd, niter, launches = run_eom_on_device(sip, case, record)
    print(f"\nreference CCSD (DIIS, {niter} iterations: ccsd_energy {e_ccsd:.14f}, golden {lw.golden_ccsd()[1]:.14f}) + EOM-CCSD programs on "
          f"the device ({case}, record={record}): roots " + ", ".join(f"{r:.14f}" for r in roots) +
          f" (goldens " + ", ".join(f"{r:.14f}" for r in g["sek0"]) + f"), {launches} launches")
    assert abs(e_ccsd - lw.golden_ccsd()[1]) < 1e-11 and niter == 15      # the run stopped at cc_conv = 1e-10, as the reference's
    for got, want in zip(roots, g["sek0"]):
        assert abs(got - want) < g["tolerance"], (roots, g["sek0"])
    assert max(abs(a - b) for a, b in zip(roots, g["sek0"])) < 2e-9
    assert launches > 0
