"""Shared driver of the energy tests (water / 3-21G: the reference's lccd_frozencore_test, eom_lccd_test, eom_mp2_test,
lccsd_test, eom_test / eom_ccsd_water_test; hydrogen fluoride / 3-21G: second_ccsdpt_test, lamccsdpt_test): inputs from
the decoded `.dat` (tests/golden/water_321g_setup.json) through oracle/qm_inputs.py (numpy integrals + RHF, test
infrastructure), then the reference's programs (tests/golden/tran_program.sialx, lccd_program.sialx, lccsd_program.sialx,
ccsd_program.sialx, and the restated ccsd_t_restated.sialx) walked block by block
by aces4_b200/sial_frontend.py on a backend -- the CPU oracle here, libsipgpu in tests/test_gpu_lccd_water_energy.py --
and the converged energy compared with the reference's golden values (test/test_qm.cpp:447-462)."""
import functools
import json
import os
import re

import numpy as np

from oracle import qm_inputs as qm

HERE = os.path.dirname(os.path.abspath(__file__))
PROGRAM = open(os.path.join(HERE, "golden", "lccd_program.sialx")).read()
PROGRAM_LCCSD = open(os.path.join(HERE, "golden", "lccsd_program.sialx")).read()
PROGRAM_CCSD = open(os.path.join(HERE, "golden", "ccsd_program.sialx")).read()
PROGRAM_PT = open(os.path.join(HERE, "golden", "ccsd_t_restated.sialx")).read()
PROGRAM_TRAN = open(os.path.join(HERE, "golden", "tran_program.sialx")).read()
FIXTURE = json.load(open(os.path.join(HERE, "golden", "water_321g_setup.json")))
GOLDEN = FIXTURE["golden"]
# array name -> index kinds of its declared dimensions
KINDS = {"ca": ("ao", "p"), "aoint": ("ao",) * 4, "vpiqj": ("p", "o", "p", "o"), "viaai": ("o", "v", "v", "o"),
         "vaaii": ("v", "v", "o", "o"), "t2old_ab": ("v", "o", "v", "o"), "t2new_ab": ("v", "o", "v", "o"),
         "tao_ab": ("ao", "o", "ao", "o"), "t2ao_ab": ("ao", "o", "ao", "o"), "tdaixj": ("v", "o", "ao", "o"),
         # LCCSD only (tests/golden/lccsd_program.sialx)
         "vspipi": ("p", "o", "p", "o"), "vaaai": ("v", "v", "v", "o"), "t2old_aa": ("v", "o", "v", "o"),
         "t1a_old": ("v", "o"), "t1a_new": ("v", "o"),
         # CCSD only (tests/golden/ccsd_program.sialx)
         "tau_ab": ("v", "o", "v", "o"), "taup_ab": ("v", "o", "v", "o"), "taup_aa": ("v", "o", "v", "o"),
         "e5aiai": ("v", "o", "v", "o"), "e5aibj": ("v", "o", "v", "o"), "e6aibj": ("v", "o", "v", "o"),
         "wiibb": ("o", "o", "v", "v"), "t1a_ax": ("v", "ao"), "fae_a": ("v", "v"), "fme_a": ("o", "v"),
         "fmi_a": ("o", "o"), "wminj_ab": ("o", "o", "o", "o"),
         # (T) only (tests/golden/ccsd_t_restated.sialx): rank-6 arrays
         "x3": ("v", "o") * 3, "w3": ("v", "o") * 3, "v3": ("v", "o") * 3,
         # integral transformation only (tests/golden/tran_program.sialx): partially transformed classes
         "vxxxi": ("ao", "ao", "ao", "o"), "vxxii": ("ao", "ao", "o", "o"), "vxixi": ("ao", "o", "ao", "o"),
         "vixxi": ("o", "ao", "ao", "o"), "vxxai": ("ao", "ao", "v", "o"), "vxipi": ("ao", "o", "p", "o"),
         "vxaii": ("ao", "v", "o", "o"), "vixai": ("o", "ao", "v", "o"), "vxaai": ("ao", "v", "v", "o"),
         "vsaaai": ("v", "v", "v", "o")}
MO_CLASSES = ("vpiqj", "vspipi", "vaaii", "viaai", "vaaai")      # what tran_program.sialx produces for the CC programs
EMPTY = ("t2old_ab", "t2new_ab", "tao_ab", "t2ao_ab", "tdaixj", "t2old_aa", "t1a_old", "t1a_new", "tau_ab", "taup_ab",
         "taup_aa", "e5aiai", "e5aibj", "e6aibj", "wiibb", "t1a_ax", "fae_a", "fme_a", "fmi_a", "wminj_ab", "x3", "w3", "v3",
         "vxxxi", "vxxii", "vxixi", "vixxi", "vxxai", "vxipi", "vxaii", "vixai", "vxaai", "vsaaai")
# cases: (setup file, segmentation).  Segmentations: the .dat's own (frozen core: moa [1 | 4 | 8], ao [11, 2]; all
# electron: moa [5 | 8], ao [13]) and a finer one of the same orbitals
FROZEN, ALL = "lccd_frozencore_test.dat", "eom_lccd_test.dat"
CASES = {"dat": (FROZEN, None), "fine": (FROZEN, {"moa": [1, 2, 2, 3, 5], "occ": (2, 3), "virt": (4, 5), "ao": [6, 5, 2]}),
         "all_dat": (ALL, None), "all_fine": (ALL, {"moa": [2, 3, 3, 5], "occ": (1, 2), "virt": (3, 4), "ao": [4, 7, 2]}),
         # hydrogen fluoride / 3-21G (the reference's second_ccsdpt_test and lamccsdpt_test): all electron, frozen core
         "hf_dat": ("second_ccsdpt_test.dat", None), "hf_fc_dat": ("lamccsdpt_test.dat", None),
         "hf_fine": ("second_ccsdpt_test.dat", {"moa": [2, 3, 2, 4], "occ": (1, 2), "virt": (3, 4), "ao": [3, 6, 2]}),
         "hf_fc_fine": ("lamccsdpt_test.dat", {"moa": [1, 3, 1, 2, 4], "occ": (2, 3), "virt": (4, 5), "ao": [3, 6, 2]}),
         # neon / cc-pVDZ with spherical d functions (the reference's DISABLED_ccsdpt_test = BASELINE config 2, ccsdpt_test.dat)
         "ne_dat": ("ccsdpt_test.dat", None),
         "ne_fine": ("ccsdpt_test.dat", {"moa": [2, 3, 4, 5], "occ": (1, 2), "virt": (3, 4), "ao": [3, 6, 5]})}


def golden(case):
    """(lccd_correlation or None, lccd_energy, mp2_energy or None) the reference's tests assert for this case"""
    if CASES[case][0] == FROZEN:
        return GOLDEN["lccd_correlation"], GOLDEN["lccd_energy"], None
    g = GOLDEN["all_electron"]
    return None, g["lccd_energy"], g["mp2_energy"]


def golden_ccsd():
    """(ccsd_energy converged to 1e-12, the same run stopped at 1e-10) of the reference's eom_test / eom_ccsd_water_test"""
    g = GOLDEN["all_electron"]
    return g["ccsd_energy"], g["ccsd_energy_cc_conv_1e-10"]


def golden_lccsd():
    """(lccsd_correlation, lccsd_energy) of the reference's lccsd_test (all electron: cases all_dat / all_fine)"""
    g = GOLDEN["all_electron"]
    return g["lccsd_correlation"], g["lccsd_energy"]


@functools.lru_cache(maxsize=None)
def scf(setup_name=FROZEN):
    """integrals + RHF of one setup; (setup, basis, S, eri, e_nuc, e_scf, eps, C)"""
    setup = FIXTURE["setups"][setup_name]
    basis = qm.basis_from_setup(setup)
    S, T, V, eri = qm.ao_integrals(basis)
    e_nuc = qm.nuclear_repulsion(basis)
    e_scf, eps, C, _ = qm.rhf(S, T + V, eri, setup["ints"]["naocc"], e_nuc)
    return setup, basis, S, eri, e_nuc, e_scf, eps, C


def inputs(case):
    """-> dict(segs, index_base, moa_seg_ranges, fock, arrays {name: {segment tuple: block}}, e_scf)"""
    setup_name, sg = CASES[case]
    setup, _, _, eri, _, e_scf, eps, C = scf(setup_name)
    if sg is None:
        it = setup["ints"]
        sg = {"moa": setup["segments"]["moa"], "occ": (it["baocc"], it["eaocc"]), "virt": (it["bavirt"], it["eavirt"]),
              "ao": setup["segments"]["ao"]}
    moa = sg["moa"]
    off = np.concatenate([[0], np.cumsum(moa)])
    occ = slice(off[sg["occ"][0] - 1], off[sg["occ"][1]])
    virt = slice(off[sg["virt"][0] - 1], off[sg["virt"][1]])
    segs = {"o": moa[sg["occ"][0] - 1: sg["occ"][1]], "v": moa[sg["virt"][0] - 1: sg["virt"][1]], "ao": sg["ao"]}
    segs["p"] = segs["o"] + segs["v"]
    dense = qm.mo_classes(eri, C, occ, virt)
    dense["aoint"] = eri
    dense["ca"] = np.hstack([C[:, occ], C[:, virt]])
    arrays = {name: qm.split_blocks(dense[name], [segs[k] for k in KINDS[name]]) for name in dense}
    for name in EMPTY:
        arrays[name] = {}
    return {"segs": segs, "index_base": {"o": sg["occ"][0] - 1, "v": sg["virt"][0] - 1}, "moa_seg_ranges": list(moa),
            "fock": np.asfortranarray(np.diag(eps)), "arrays": arrays, "e_scf": e_scf}


def converge(walker, value, tol=1e-12, max_iter=80):
    """main program (starting guess + second-order energy), then `proc iteration` until the energy is stationary;
    `value` turns a walker scalar into a float.  -> (mp2 energy, [energy per iteration])"""
    e_mp2 = value(walker.run()["ecorrab"])
    hist = []
    for _ in range(max_iter):
        hist.append(value(walker.run_proc("iteration")["ecorrab"]))
        if len(hist) > 1 and abs(hist[-1] - hist[-2]) < tol:
            return e_mp2, hist
    raise AssertionError(f"LCCD iterations did not converge: {hist[-3:]}")


PERSISTED = ("VSpipi", "Vaaii", "Viaai", "Vaaai", "Vpiqj")


def chained_through_persistence(tran_text, cc_text):
    """the two programs chained the way the reference chains them (tran_rhf_no4v.sialx:632-637 `set_persistent X "X"` at
    the end of the transformation; rlccd_rhf.sialx:243-251 / rccsd_rhf.sialx:225-244 `PROC READ_2EL` with
    `restore_persistent X "X"` at the start of the CC program): -> (transformation text, CC text)"""
    tail = "".join(f'set_persistent {x} "{x}"\n' for x in PERSISTED)
    tran = tran_text.replace("endsial tran_program", tail + "endsial tran_program")
    assert tran != tran_text
    declared = [x for x in PERSISTED if re.search(rf"^served {x}\[", cc_text, re.M)]     # LCCD reads four of the five
    head = "proc read_2el\n" + "".join(f'restore_persistent {x} "{x}"\n' for x in declared) + "server_barrier\nendproc read_2el\n\n"
    i = cc_text.index("proc iguess")
    cc = cc_text[:i] + head + cc_text[i:]
    j = cc.index("call iguess")
    cc = cc[:j] + "call read_2el\n" + cc[j:]
    return tran, cc


def dense_ca(inp):
    """the whole static array ca[mu,p] (norb x active MOs) and the segment extents of its two dimensions"""
    seg = [inp["segs"]["ao"], inp["segs"]["p"]]
    return qm.join_blocks(inp["arrays"]["ca"], seg), seg


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own (T) programs (tests/golden/rccsdpt_aaa_program.sialx / rccsdpt_aab_program.sialx, generated from
# src/sialx/qm/cc/rccsdpt_aaa.sialx / rccsdpt_aab.sialx by scripts/make_ccsdpt_aab_golden.py)
PROGRAM_PT_AAA = open(os.path.join(HERE, "golden", "rccsdpt_aaa_program.sialx")).read()
PROGRAM_PT_AAB = open(os.path.join(HERE, "golden", "rccsdpt_aab_program.sialx")).read()
# VSaaai[a2,a,a1,i] = Vaaai[a2,a,a1,i] - Vaaai[a1,a,a2,i]: the last step of TRAN_TRAN4 (rccsdpt_aab.sialx:431-439), whose
# transformation procedures the generated programs replace by restore_persistent of their results
VSAAAI_FRAGMENT = """
moaindex a = bavirt: eavirt
moaindex a1 = bavirt: eavirt
moaindex a2 = bavirt: eavirt
moaindex i = baocc: eaocc
served Vaaai[a,a1,a2,i]
served VSaaai[a,a1,a2,i]
temp t[a,a1,a2,i]
temp t1[a,a1,a2,i]
pardo a2, a, a1, i
   request Vaaai[a2,a,a1,i]
   request Vaaai[a1,a,a2,i]
   t[a2,a,a1,i]  = Vaaai[a2,a,a1,i]
   t1[a2,a,a1,i] = Vaaai[a1,a,a2,i]
   t[a2,a,a1,i] -= t1[a2,a,a1,i]
   prepare VSaaai[a2,a,a1,i] = t[a2,a,a1,i]
endpardo a2, a, a1, i
server_barrier
"""
# persistence labels the (T) programs restore -> array names of the CCSD / transformation programs that produce them
PT_LABELS = {"t1a_old": "t1a_old", "T2old_aa": "t2old_aa", "T2old_ab": "t2old_ab", "Vpiqj": "vpiqj", "VSpipi": "vspipi",
             "Vaaai": "vaaai", "VSaaai": "vsaaai"}
# label -> name of the array that holds it inside the (T) programs (what to persist again for the program that follows)
PT_HOLDERS = {"t1a_old": "t1a_old", "T2old_aa": "tsaiai", "T2old_ab": "t2aiai", "Vpiqj": "vpiqj", "VSpipi": "vspipi",
              "Vaaai": "vaaai", "VSaaai": "vsaaai"}


def pt_constants(inp):
    """predefined ints the (T) programs read: naocc = number of occupied ORBITALS (the range of the simple indices ii, jj),
    baocc / eaocc = first / last occupied segment of moa_seg_ranges"""
    return {"naocc": sum(inp["segs"]["o"]), "baocc": inp["index_base"]["o"] + 1,
            "eaocc": inp["index_base"]["o"] + len(inp["segs"]["o"])}


def pt_array_kinds(program):
    """index kinds of every served / distributed array a (T) program declares ('s' = simple index: blocks of extent 1)"""
    return {n: tuple(program.index_kind[d] for d in decl) for n, (k, decl) in program.arrays.items() if k in ("served", "distributed")}


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 3: the reference's right-hand EOM-CCSD program (tests/golden/eom_ccsd_right_program.sialx, generated from
# src/sialx/qm/eom/eom_ccsd_rhf_right.sialx + eom_rhf_hbar.sialx + eom_rhf_vars.sialx + eom_rhf_defs.sialx by
# scripts/make_eom_golden.py) on water / 3-21G (test/eom_ccsd_water_test.dat)
PROGRAM_EOM = open(os.path.join(HERE, "golden", "eom_ccsd_right_program.sialx")).read()
PROGRAM_EOM_FULL = open(os.path.join(HERE, "golden", "eom_ccsd_right_full_program.sialx")).read()   # the same, whole file
PROGRAM_RCIS_D = open(os.path.join(HERE, "golden", "rcis_d_rhf_program.sialx")).read()             # src/sialx/qm/eom/rcis_d_rhf.sialx
CASES["cis_dat"] = ("cis_test.dat", None)           # hydrogen fluoride / 3-21G, two roots (the reference's cis_test)
PROGRAM_EOM_LEFT = open(os.path.join(HERE, "golden", "eom_ccsd_left_program.sialx")).read()
EOM_SETUP = "eom_ccsd_water_test.dat"
CASES["eom_dat"] = (EOM_SETUP, None)
CASES["eom_fine"] = (EOM_SETUP, {"moa": [2, 3, 3, 5], "occ": (1, 2), "virt": (3, 4), "ao": [4, 7, 2]})
# persistence labels READ_AMP restores (eom_ccsd_rhf_right.sialx:13-29) -> arrays of the CCSD / transformation programs
EOM_LABELS = {"t1a_old": "t1a_old", "T2old_aa": "t2old_aa", "T2old_ab": "t2old_ab", "VSpipi": "vspipi", "Vaaii": "vaaii",
              "Viaai": "viaai", "Vaaai": "vaaai", "VSaaai": "vsaaai", "Vpiqj": "vpiqj"}


def eom_constants():
    """the predefined ints / scalars of the setup file the EOM program reads: eom_roots, eom_tol, cc_iter (the range of the
    subspace indices), baocc ..."""
    setup = FIXTURE["setups"][EOM_SETUP]
    return {**setup["ints"], **setup["scalars"]}


def cis_guess(inp, dense):
    """The starting vectors C1_a[kstate,a,i] of the EOM program = the converged singlet CIS vectors of rcis_rhf.sialx, which
    runs right before it.  Here: dense diagonalisation of the CIS matrix (oracle/qm_inputs.py, test infrastructure) -- pinned
    by the reference's CIS goldens of the same molecule (DISABLED_eom_test, test/test_qm.cpp:265-272).
    dense: name -> dense MO integral class.  -> (CIS energies, {(kstate, virtual segment, occupied segment): block [1,a,i]})"""
    eps = np.diag(inp["fock"])
    no = sum(inp["segs"]["o"])
    nroots = FIXTURE["setups"][EOM_SETUP]["ints"]["eom_roots"]
    e, vec = qm.cis_singlets(eps[:no], eps[no:], dense["vpiqj"][no:, :, no:, :], dense["vaaii"], nroots)
    blocks = {}
    for k in range(nroots):
        for (sa, si), b in qm.split_blocks(vec[k], [inp["segs"]["v"], inp["segs"]["o"]]).items():
            blocks[(k + 1, sa, si)] = np.asfortranarray(b[None, :, :])
    return e, blocks


def eom_array_kinds(program):
    """index kinds of every served / distributed array the EOM program declares ('s': simple index, blocks of extent 1)"""
    return {n: tuple(program.index_kind[d] for d in decl) for n, (k, decl) in program.arrays.items() if k in ("served", "distributed")}


def eom_simple_extents(program, constants):
    """number of values of every simple index range that occurs as an array dimension: {declared label: count}"""
    out = {}
    for lab, (lo, hi) in program.simple_range.items():
        lo, hi = (int(x) if x.isdigit() else int(constants[x]) for x in (lo, hi))
        out[lab] = hi - lo + 1
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CCSD program VERBATIM (tests/golden/rccsd_rhf_program.sialx = src/sialx/qm/cc/rccsd_rhf.sialx, generated by
# scripts/make_cc_program_goldens.py): the `DO KITER` loop with DIIS and the convergence test at the setup's cc_conv
PROGRAM_RCCSD = open(os.path.join(HERE, "golden", "rccsd_rhf_program.sialx")).read()
PROGRAM_RLCCD = open(os.path.join(HERE, "golden", "rlccd_rhf_program.sialx")).read()
PROGRAM_RLCCSD = open(os.path.join(HERE, "golden", "rlccsd_rhf_program.sialx")).read()
PROGRAM_RLAMBDA = open(os.path.join(HERE, "golden", "rlambda_rhf_program.sialx")).read()   # src/sialx/qm/cc/rlambda_rhf.sialx
CASES["hf_fc_virt_fine"] = ("lamccsdpt_test.dat", {"moa": [1, 4, 2, 4], "occ": (2, 2), "virt": (3, 4), "ao": [3, 6, 2]})
# two ACTIVE occupied segments of two orbitals each (hf_fc_fine above cuts them 3 + 1: a one-orbital segment, which the reference's
# energy_denominator_rhf.F takes for a simple index)
CASES["hf_fc_occ22"] = ("lamccsdpt_test.dat", {"moa": [1, 2, 2, 2, 4], "occ": (2, 3), "virt": (4, 5), "ao": [3, 6, 2]})
CASES["lam_dat"] = ("rlambda_test.dat", None)       # hydrogen fluoride / 3-21G, cc_conv 1e-12 (the reference's rlambda_test)
CASES["lam_fine"] = ("rlambda_test.dat", {"moa": [2, 3, 2, 4], "occ": (1, 2), "virt": (3, 4), "ao": [3, 6, 2]})
PROGRAM_RLAMPT_AAA = open(os.path.join(HERE, "golden", "rlamccsdpt_aaa_program.sialx")).read()   # src/sialx/qm/cc/rlamccsdpt_aaa.sialx
PROGRAM_RLAMPT_AAB = open(os.path.join(HERE, "golden", "rlamccsdpt_aab_program.sialx")).read()   # src/sialx/qm/cc/rlamccsdpt_aab.sialx
PROGRAM_RCIS = open(os.path.join(HERE, "golden", "rcis_rhf_program.sialx")).read()     # src/sialx/qm/eom/rcis_rhf.sialx
PROGRAM_TRAN_NO4V = open(os.path.join(HERE, "golden", "tran_rhf_no4v_program.sialx")).read()    # src/sialx/qm/utility/tran_rhf_no4v.sialx


def setup_constants(case):
    """the predefined ints / scalars of the case's setup file (cc_conv, cc_iter, scf_hist, baocc ...)"""
    setup = FIXTURE["setups"][CASES[case][0]]
    out = {**setup["ints"], **setup["scalars"]}
    sg = CASES[case][1]
    if sg is not None:          # a finer segmentation of the same orbitals: the segment ranges of the active spaces move with it
        out.update(baocc=sg["occ"][0], eaocc=sg["occ"][1], bavirt=sg["virt"][0], eavirt=sg["virt"][1], norb=len(sg["ao"]))
    return out


def all_orbital_statics(case, inp):
    """the static arrays of the defs files, declared over ALL molecular orbital segments (`moaindex aces_defs_pa = 1: eavirt`):
    ca[mu,pa] with the frozen-core columns in place, fock_a[pa,pa] -> {label: {segment tuple: block}}"""
    C = scf(CASES[case][0])[7]
    moa = inp["moa_seg_ranges"]
    return {"ca": qm.split_blocks(C, [inp["segs"]["ao"], moa]), "fock_a": qm.split_blocks(inp["fock"], [moa, moa])}


def segs_with_all_orbitals(inp):
    """segment extents per index kind, plus 'pa' = every molecular orbital segment of the setup (frozen core included)"""
    return {**inp["segs"], "pa": list(inp["moa_seg_ranges"])}


def program_array_kinds(program):
    """index kinds of every served / distributed array a program declares ('s': simple index, blocks of extent 1)"""
    return {n: tuple(program.index_kind[d] for d in decl) for n, (k, decl) in program.arrays.items() if k in ("served", "distributed")}


def used_arrays(text):
    """names of the arrays a program text moves blocks of (request / get / put / prepare / restore_persistent / set_persistent)"""
    return {n.lower() for n in re.findall(r"(?im)^\s*(?:request|get|put|prepare|restore_persistent|set_persistent)\s+([a-z_]\w*)", text)}


def device_program_arrays(sip, program, text, constants, segs, skip=()):
    """one (zero-filled) api.DistArray per served / distributed array the program text touches; a simple-index dimension
    becomes as many one-element segments as the index has values"""
    seg_ext = dict(segs)
    seg_ext["p"] = list(segs["o"]) + list(segs["v"])
    simple = eom_simple_extents(program, constants)
    used = used_arrays(text)
    out = {}
    for name, (kind, decl) in program.arrays.items():
        if kind not in ("served", "distributed") or name not in used or name in skip:
            continue
        out[name] = sip.DistArray([[1] * simple[d] if program.index_kind[d] == "s" else seg_ext[program.index_kind[d]] for d in decl])
        out[name].fill_local(0.0)
    return out


def restored_labels(text):
    """[(array name, label)] of the `restore_persistent` statements of a program text.  Served arrays persisted by an earlier
    program live in the servers' files and outlive a restore (disk_backed_block_map.restore_persistent_array), so a later program
    can restore a label again although the program in between did not `set_persistent` it: the harnesses hand the arrays a
    program restored over again under the same labels before the next program runs."""
    return [(n.lower(), lab) for n, lab in re.findall(r'(?im)^\s*restore_persistent\s+(\w+)\s+"(\w+)"', text)]


@functools.lru_cache(maxsize=None)
def dipole_data(setup_name):
    """what the dipole integral engine (`compute_dipole_integrals`, OED package: out of scope) and the SCF program would supply:
    {"dipole_integrals": <mu|r|nu> [3,nao,nao], "nuclear_dipole": [3]} for Walker(host_data=...) and the SCF dipole moment
    (persistence label "scf_dipole"), from oracle/qm_inputs.py.  Pinned by the reference's SCF dipole of rlambda_test
    (test/test_qm.cpp:317: 0.84792717246707)."""
    setup, basis, _, _, _, _, _, C = scf(setup_name)
    D = qm.dipole_integrals(basis)
    Z, X = np.array(basis["charge"]), np.array(basis["coords"])
    nuc = (Z[:, None] * X).sum(0)
    nocc = setup["ints"]["naocc"]
    P = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    scf_dipole = nuc - np.einsum("dmn,mn->d", D, P)
    return {"dipole_integrals": D, "nuclear_dipole": nuc}, {(k + 1,): float(scf_dipole[k]) for k in range(3)}
