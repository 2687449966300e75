"""BASELINE config 3 ("test/eom_ccsd_water_test.dat EOM-CCSD water, Davidson sigma-vector contractions") ON THE DEVICE
through the reference's own right-hand EOM-CCSD program: tests/golden/eom_ccsd_right_program.sialx (= src/sialx/qm/eom/
eom_ccsd_rhf_right.sialx + eom_rhf_hbar.sialx, see scripts/make_eom_golden.py) walked by the SIAL front-end on libsipgpu
after the reference's CCSD program: H-bar, its diagonal, and per Davidson step the sigma-vector contractions on rank-5
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
    from aces4_b200.sial_frontend import DeviceBackend, Program, Walker

    inp = lw.inputs(case)
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    arrays = {}
    for name, kinds in lw.KINDS.items():
        A = sip.DistArray([inp["segs"][k] for k in kinds])
        A.fill_local(0.0)
        upload(sip, A, inp["arrays"][name])
        arrays[name] = A
    sip.sync()
    be = DeviceBackend(sip, arrays, record=record)
    be.fock = sip.DeviceBlock.from_numpy(inp["fock"])
    be.seg_ranges = inp["moa_seg_ranges"]
    w = Walker(Program(lw.PROGRAM_CCSD), be, inp["segs"], index_base=inp["index_base"])
    _, hist = lw.converge(w, be.value, tol=1e-12, max_iter=150)
    Walker(Program(lw.VSAAAI_FRAGMENT), be, inp["segs"], index_base=inp["index_base"]).run()
    sip.sync()
    dense = {n: qm.join_blocks({idx: arrays[n].block_view(idx).to_numpy() for idx in inp["arrays"][n]},
                               [inp["segs"][k] for k in lw.KINDS[n]]) for n in ("vpiqj", "vaaii")}
    e_cis, c1 = lw.cis_guess(inp, dense)

    prog = Program(lw.PROGRAM_EOM)
    consts = lw.eom_constants()
    seg_ext = dict(inp["segs"])
    seg_ext["p"] = list(inp["segs"]["o"]) + list(inp["segs"]["v"])
    simple = lw.eom_simple_extents(prog, consts)
    used = {n.lower() for n in __import__("re").findall(r"(?im)^\s*(?:request|get|put|prepare|restore_persistent)\s+([a-z_]\w*)", lw.PROGRAM_EOM)}
    parr = {}
    for name, (kind, decl) in prog.arrays.items():
        if kind not in ("served", "distributed") or name not in used or name == "aoint":
            continue
        parr[name] = sip.DistArray([[1] * simple[d] if prog.index_kind[d] == "s" else seg_ext[prog.index_kind[d]] for d in decl])
        parr[name].fill_local(0.0)
    upload(sip, parr["c1_a"], c1)
    sip.sync()
    for label, arr in lw.EOM_LABELS.items():
        arrays[arr].persist(label)
    parr["c1_a"].persist("C1_a")
    parr["aoint"] = arrays["aoint"]
    fock_a = sip.DistArray([seg_ext["p"], seg_ext["p"]])
    fock_a.fill_local(0.0)
    upload(sip, fock_a, qm.split_blocks(inp["fock"], [seg_ext["p"], seg_ext["p"]]))
    parr["ca"], parr["fock_a"] = arrays["ca"], fock_a
    be2 = DeviceBackend(sip, parr, record=record)
    be2.fock, be2.seg_ranges = be.fock, inp["moa_seg_ranges"]
    l0 = sip.kernel_launches()
    w2 = Walker(prog, be2, inp["segs"], index_base=inp["index_base"], constants=consts)
    w2.run()
    roots = [w2.tables["sek0"][(k,)] for k in range(1, len(e_cis) + 1)]
    return roots, e_cis, inp["e_scf"] + hist[-1], sip.kernel_launches() - l0


@pytest.mark.timeout(1500, method="thread")
@pytest.mark.parametrize("case,record", [("eom_dat", True), ("eom_fine", False)])
def test_reference_eom_program_on_the_device(sip, case, record):
    g = lw.GOLDEN["eom_ccsd_water_test"]
    roots, e_cis, e_ccsd, launches = run_eom_on_device(sip, case, record)
    print(f"\nreference EOM-CCSD program on the device ({case}, record={record}): roots " + ", ".join(f"{r:.14f}" for r in roots) +
          f" (goldens " + ", ".join(f"{r:.14f}" for r in g["sek0"]) + f"), {launches} launches")
    assert abs(e_ccsd - lw.golden_ccsd()[0]) < 1e-10
    for got, want in zip(roots, g["sek0"]):
        assert abs(got - want) < g["tolerance"], (roots, g["sek0"])
    assert max(abs(a - b) for a, b in zip(roots, g["sek0"])) < 2e-9
    assert abs(roots[0] - lw.GOLDEN["eom_test"]["eom_sek0"][0]) < 1e-10
    assert launches > 0
