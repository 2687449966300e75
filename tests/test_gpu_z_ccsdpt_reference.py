"""BASELINE config 2 ("test/ccsdpt_test.dat CCSD(T) on 1 B200, full (T) triples contraction path") ON THE DEVICE through
the reference's own triples programs: tests/golden/rccsdpt_aaa_program.sialx / rccsdpt_aab_program.sialx (=
src/sialx/qm/cc/rccsdpt_aaa.sialx / rccsdpt_aab.sialx, see scripts/make_ccsdpt_aab_golden.py) walked by the SIAL front-end
on libsipgpu after the reference's CCSD program: every stripi slice, permute, one-segment contraction, `+=` into the rank-6
local blocks, energy_denominator_rhf on [a,ii,a1,jj,b,k1] blocks, scalar contraction and put / get of the simple-index
arrays is a C-ABI call.  Goldens: the four spin components of second_ccsdpt_test (test/test_qm.cpp:110-124) at the
reference's 1e-10, eaab / esaab of ccsdpt_test.dat (:45-48) at north_star's 1e-9.  CPU twin (oracle backend):
tests/test_ccsdpt_reference_programs_cpu.py."""
import re

import numpy as np
import pytest

import lccd_water as lw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


def device_arrays(sip, inp):
    out = {}
    for name, kinds in lw.KINDS.items():
        A = sip.DistArray([inp["segs"][k] for k in kinds])
        A.fill_local(0.0)
        for idx, b in inp["arrays"][name].items():
            view = A.block_view(idx)
            sip._check(sip.lib().sipgpu_h2d(view.ptr, sip._hp(np.asfortranarray(b)), view.size), "h2d")
        out[name] = A
    sip.sync()
    return out


def run_pt_on_device(sip, case, which, record):
    from aces4_b200.sial_frontend import DeviceBackend, Program, Walker

    inp = lw.inputs(case)
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    arrays = device_arrays(sip, inp)
    be = DeviceBackend(sip, arrays, record=record)
    be.fock = sip.DeviceBlock.from_numpy(inp["fock"])
    be.seg_ranges = inp["moa_seg_ranges"]
    w = Walker(Program(lw.PROGRAM_CCSD), be, inp["segs"], index_base=inp["index_base"])
    _, hist = lw.converge(w, be.value, max_iter=150)
    Walker(Program(lw.VSAAAI_FRAGMENT), be, inp["segs"], index_base=inp["index_base"]).run()
    consts = lw.pt_constants(inp)
    seg_ext = dict(inp["segs"])
    seg_ext["p"] = list(inp["segs"]["o"]) + list(inp["segs"]["v"])
    seg_ext["s"] = [1] * consts["naocc"]
    holders = {label: arrays[arr] for label, arr in lw.PT_LABELS.items()}     # who holds each persistence label right now
    out, launches = {}, 0
    for name in which:
        for label, A in holders.items():
            A.persist(label)
        sip.persist_scalar("totenerg", 0.0)
        sip.persist_scalar("ccsd_energy", 0.0)
        prog = Program(lw.PROGRAM_PT_AAA if name == "aaa" else lw.PROGRAM_PT_AAB)
        used = {n.lower() for n in re.findall(r"(?i)\b(?:request|get|put|prepare|restore_persistent)\s+([a-z_]\w*)",
                                                           lw.PROGRAM_PT_AAA if name == "aaa" else lw.PROGRAM_PT_AAB)}
        parr = {n: sip.DistArray([seg_ext[k] for k in kinds]) for n, kinds in lw.pt_array_kinds(prog).items()
                if n in used and all(k in seg_ext for k in kinds)}
        for A in parr.values():
            A.fill_local(0.0)
        be2 = DeviceBackend(sip, parr, record=record)
        be2.fock, be2.seg_ranges = be.fock, inp["moa_seg_ranges"]
        l0 = sip.kernel_launches()
        sc = Walker(prog, be2, inp["segs"], index_base=inp["index_base"], constants=consts).run()
        launches += sip.kernel_launches() - l0
        out.update({k: be2.value(sc[k]) for k in (("eaaa", "esaaa") if name == "aaa" else ("eaab", "esaab"))})
        restored = {label: parr[h] for label, h in lw.PT_HOLDERS.items() if h in parr and label in
                    ("t1a_old", "T2old_aa", "T2old_ab", "VSpipi", "VSaaai") + (("Vpiqj", "Vaaai") if name == "aab" else ())}
        holders = {label: restored.get(label, A) for label, A in holders.items()}
        for label, A in holders.items():     # labels this program did not restore are still registered: take them back
            if label not in restored:
                A.restore(label)
    return out, hist, launches


@pytest.mark.timeout(900, method="thread")
@pytest.mark.parametrize("case,record", [("hf_dat", True), ("hf_virt_fine", False)])
def test_reference_triples_programs_on_the_device(sip, case, record):
    lw.CASES["hf_virt_fine"] = ("second_ccsdpt_test.dat", {"moa": [5, 2, 4], "occ": (1, 1), "virt": (2, 3), "ao": [3, 6, 2]})
    g = lw.GOLDEN["hf"]
    got, hist, launches = run_pt_on_device(sip, case, ("aaa", "aab"), record)
    print(f"\\nreference (T) programs on the device ({case}, record={record}): " +
          ", ".join(f"{k} {v:.14e} (golden {g[k]:.8e})" for k, v in got.items()) + f", {launches} launches")
    for name in ("eaaa", "esaaa", "eaab", "esaab"):
        assert abs(got[name] - g[name]) < lw.GOLDEN["tolerance"], (name, got[name], g[name])
    assert abs(sum(got.values()) - (g["ccsdpt_energy"] - g["ccsd_energy"])) < 1e-10
    assert launches > 0


@pytest.mark.timeout(900, method="thread")
def test_reference_aab_program_on_the_device_matches_ccsdpt_test_dat(sip):
    g = lw.GOLDEN["ne_ccsdpt_test"]
    got, hist, launches = run_pt_on_device(sip, "ne_dat", ("aab",), True)
    print(f"\\nccsdpt_test.dat through rccsdpt_aab on the device: eaab {got['eaab']:.16f} (golden {g['eaab']:.16f}), "
          f"esaab {got['esaab']:.16e} (golden {g['esaab']:.16e})")
    assert abs(got["eaab"] - g["eaab"]) < 1e-9 and abs(got["esaab"] - g["esaab"]) < 1e-9
    assert abs(got["eaab"] - (-0.0010909776279972)) < 1e-12 and abs(got["esaab"] - 8.554806688752e-05) < 1e-12
