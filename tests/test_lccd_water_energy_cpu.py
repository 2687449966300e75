"""Energy-level parity on the CPU side (no GPU): the reference's golden energies of lccd_frozencore_test
(test/test_qm.cpp:431-468: water, 3-21G, drop_mo=1-1) and of the all-electron runs of the same molecule (:639-678 LCCD,
:720-759 MP2) reproduced by (1) the numpy input stage of oracle/qm_inputs.py
(SCF energy: pins the integrals, i.e. the INPUTS of the hot path) and (2) the reference's LCCD amplitude equations
(tests/golden/lccd_program.sialx) walked by the SIAL front-end on the ORACLE backend (correlation energy: pins the
front-end's call stream and the oracle's block arithmetic at the energy level).  The same program on libsipgpu:
tests/test_gpu_lccd_water_energy.py."""
import numpy as np
import pytest

import lccd_water as lw
from aces4_b200.sial_frontend import Program, Walker
from oracle import qm_inputs as qm
from sial_oracle_backend import OracleBackend


def test_boys_function_against_quadrature():
    T = np.array([0.0, 1e-9, 1e-3, 0.5, 3.0, 12.0, 34.9, 35.0, 36.0, 120.0, 900.0])
    x, w = np.polynomial.legendre.leggauss(400)
    t = 0.5 * (x + 1.0)
    F = qm.boys(4, T)
    for n in range(5):
        want = np.array([0.5 * np.sum(w * t ** (2 * n) * np.exp(-Ti * t * t)) for Ti in T])
        assert np.max(np.abs(F[n] - want) / want) < 5e-13, n


@pytest.mark.parametrize("setup_name", [lw.FROZEN, lw.ALL])
def test_inputs_reproduce_the_reference_scf_energy(setup_name):
    setup, basis, S, eri, e_nuc, e_scf, eps, C = lw.scf(setup_name)
    assert abs(e_nuc - setup["scalars"]["nn_repulsion"]) < 1e-12          # geometry decoded as the reference reads it
    assert np.allclose(C.T @ S @ C, np.eye(len(S)), atol=1e-10)
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.max(np.abs(eri - eri.transpose(perm))) < 1e-13
    assert abs(e_scf - lw.GOLDEN["scf_energy"]) < lw.GOLDEN["tolerance"]  # -75.58432674274046 @ 1e-10
    assert abs(e_scf - lw.GOLDEN["all_electron"]["scf_energy"]) < lw.GOLDEN["tolerance"]   # the other tests' rounding of it


def dense_lccd(inp, tol=1e-13):
    """the same equations as dense einsums (independent of the walker and of the block arithmetic)"""
    segs, A = inp["segs"], inp["arrays"]
    j = lambda n: qm.join_blocks(A[n], [segs[k] for k in lw.KINDS[n]])  # noqa: E731
    Vp, Via, Vaa, ca, ao = j("vpiqj"), j("viaai"), j("vaaii"), j("ca"), j("aoint")
    no = sum(segs["o"])
    V, Vo, cv = Vp[no:, :, no:, :], Vp[:no, :, :no, :], ca[:, no:]
    eps = np.diag(inp["fock"])
    lo = sum(inp["moa_seg_ranges"][: inp["index_base"]["o"]])
    eo, ev = eps[lo: lo + no], eps[lo + no:]
    D = eo[None, :, None, None] + eo[None, None, None, :] - ev[:, None, None, None] - ev[None, None, :, None]
    sym = lambda X: X + X.transpose(2, 3, 0, 1)  # noqa: E731
    en = lambda T: np.einsum("aibj,aibj->", T, 2.0 * V - V.transpose(0, 3, 2, 1))  # noqa: E731
    T = 0.5 * sym(V) / D
    e_mp2, e_old = en(T), 0.0
    for _ in range(100):
        new = sym(0.5 * V) + np.einsum("akbl,ikjl->aibj", T, Vo)
        TY = np.einsum("iack->aick", Via) - np.einsum("caik->aick", Vaa)
        new += sym(np.einsum("aick,ckbj->aibj", TY, T))
        W = np.einsum("ckai->ckia", T) - np.einsum("ciak->ckia", T)
        new += sym(np.einsum("ckia,iabj->ckbj", W, Via))
        new += sym(-np.einsum("akcj,bcki->aibj", T, Vaa))
        tao = np.einsum("aibj,ma,nb->minj", T, cv, cv)
        new += np.einsum("minj,ma,nb->aibj", np.einsum("lmsn,lisj->minj", ao, tao), cv, cv)
        T = 0.5 * sym(new) / D
        e = en(T)
        if abs(e - e_old) < tol:
            return e_mp2, e
        e_old = e
    raise AssertionError("dense LCCD did not converge")


@pytest.mark.parametrize("case", ["dat", "fine", "all_dat", "all_fine"])
def test_lccd_energy_on_the_oracle_backend_matches_the_reference_golden(oracle, case):
    inp = lw.inputs(case)
    g_corr, g_total, g_mp2 = lw.golden(case)
    tol = lw.GOLDEN["tolerance"]
    e_mp2_dense, e_dense = dense_lccd(inp)
    assert abs(e_dense + inp["e_scf"] - g_total) < tol
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(Program(lw.PROGRAM), be, inp["segs"], index_base=inp["index_base"])
    e_mp2, hist = lw.converge(w, be.value)
    assert abs(e_mp2 - e_mp2_dense) < 1e-12
    e_corr = hist[-1]
    if g_corr is not None:
        assert abs(e_corr - g_corr) < tol                     # frozen core: lccd_correlation -0.12610179886435 @ 1e-10
    assert abs(e_corr + inp["e_scf"] - g_total) < tol         # lccd_energy -75.71042854160481 / -75.71210049055006 @ 1e-10
    if g_mp2 is not None:
        assert abs(e_mp2 + inp["e_scf"] - g_mp2) < tol        # all electron: mp2_energy -75.70540831822183 @ 1e-10
    assert abs(e_corr - e_dense) < 1e-11
    assert be.calls > 1000 * len(hist) or case.endswith("dat")   # the fine segmentations really are block-wise


@pytest.mark.parametrize("world", [2, 3])
def test_lccd_iteration_partitions_over_workers(oracle, world):
    """every worker walks the program and executes its share of the pardo iterations (loop_manager.cpp:468-499);
    accumulating into shared arrays, one residual evaluation by `world` workers equals the single-worker one"""
    inpw = lw.inputs("fine")
    prog = Program(lw.PROGRAM)
    bes = [OracleBackend(oracle, inpw["arrays"], fock=inpw["fock"], moa_seg_ranges=inpw["moa_seg_ranges"])
           for _ in range(world)]
    ws = [Walker(prog, bes[r], inpw["segs"], rank=r, world=world, index_base=inpw["index_base"]) for r in range(world)]
    # barrier-synchronous execution: every worker finishes a procedure before any starts the next one
    for name in ("iguess", "t2new_zero", "t2newab", "hhladder_ab", "phladder_ab"):
        for w in ws:
            w.run_proc(name)
    # reference state after the same procedures on one worker
    inp2 = lw.inputs("fine")
    be2 = OracleBackend(oracle, inp2["arrays"], fock=inp2["fock"], moa_seg_ranges=inp2["moa_seg_ranges"])
    w2 = Walker(prog, be2, inp2["segs"], index_base=inp2["index_base"])
    for name in ("iguess", "t2new_zero", "t2newab", "hhladder_ab", "phladder_ab"):
        w2.run_proc(name)
    for idx, b in inp2["arrays"]["t2old_ab"].items():
        assert np.array_equal(inpw["arrays"]["t2old_ab"][idx], b)
    for idx, b in inp2["arrays"]["t2new_ab"].items():
        assert np.max(np.abs(inpw["arrays"]["t2new_ab"][idx] - b)) <= 1e-13
    assert min(be.calls for be in bes) > 0


# ---------------------------------------------------------------------------------------------------------------------
# LCCSD (tests/golden/lccsd_program.sialx = src/sialx/qm/cc/rlccsd_rhf.sialx): singles + doubles, local arrays
# ---------------------------------------------------------------------------------------------------------------------
def dense_lccsd(inp, tol=1e-13):
    """the LCCSD equations of rlccsd_rhf.sialx as dense einsums (independent of the walker / block arithmetic)"""
    segs, A = inp["segs"], inp["arrays"]
    j = lambda n: qm.join_blocks(A[n], [segs[k] for k in lw.KINDS[n]])  # noqa: E731
    Vp, VSp, Via, Vaa, Vaaai, ca, ao = j("vpiqj"), j("vspipi"), j("viaai"), j("vaaii"), j("vaaai"), j("ca"), j("aoint")
    no = sum(segs["o"])
    V, Vo, Voovo, VS, cv = Vp[no:, :, no:, :], Vp[:no, :, :no, :], Vp[:no, :, no:, :], VSp[no:, :, :no, :], ca[:, no:]
    eps = np.diag(inp["fock"])
    lo = sum(inp["moa_seg_ranges"][: inp["index_base"]["o"]])
    eo, ev = eps[lo: lo + no], eps[lo + no:]
    D2 = eo[None, :, None, None] + eo[None, None, None, :] - ev[:, None, None, None] - ev[None, None, :, None]
    D1 = eo[None, :] - ev[:, None]
    sym = lambda X: X + X.transpose(2, 3, 0, 1)  # noqa: E731
    en = lambda T: np.einsum("aibj,aibj->", T, 2.0 * V - V.transpose(0, 3, 2, 1))  # noqa: E731
    T2, t1 = 0.5 * sym(V) / D2, np.zeros((len(ev), no))
    e_old = 0.0
    for _ in range(200):
        Taa = T2 - T2.transpose(0, 3, 2, 1)
        n1 = np.einsum("iabj,bj->ai", Via, t1) - np.einsum("acki,ck->ai", Vaa, t1) + np.einsum("kcai,ck->ai", Via, t1)
        n1 += -0.5 * np.einsum("dack,cidk->ai", Vaaai - Vaaai.transpose(2, 1, 0, 3), Taa)
        n1 += -0.5 * np.einsum("clik,akcl->ai", VS, Taa)
        n1 += np.einsum("cabj,cibj->ai", Vaaai, T2) - np.einsum("ikbj,akbj->ai", Voovo, T2)
        new = sym(0.5 * V - np.einsum("kibj,ak->aibj", Voovo, t1) + np.einsum("acbj,ci->aibj", Vaaai, t1))
        new += np.einsum("akbl,kilj->aibj", T2, Vo)
        TY = np.einsum("iack->aick", Via) - np.einsum("caik->aick", Vaa)
        new += sym(np.einsum("aick,ckbj->aibj", TY, T2))
        W = np.einsum("ckai->ckia", T2) - np.einsum("ciak->ckia", T2)
        new += sym(np.einsum("ckia,iabj->ckbj", W, Via))
        new += sym(-np.einsum("akcj,bcki->aibj", T2, Vaa))
        tao = np.einsum("aibj,ma,nb->minj", T2, cv, cv)
        new += np.einsum("minj,ma,nb->aibj", np.einsum("lmsn,lisj->minj", ao, tao), cv, cv)
        t1, T2 = n1 / D1, 0.5 * sym(new) / D2
        e = en(T2)
        if abs(e - e_old) < tol:
            return e, t1
        e_old = e
    raise AssertionError("dense LCCSD did not converge")


@pytest.mark.parametrize("case", ["all_dat", "all_fine"])
def test_lccsd_energy_on_the_oracle_backend_matches_the_reference_golden(oracle, case):
    inp = lw.inputs(case)
    g_corr, g_total = lw.golden_lccsd()
    tol = lw.GOLDEN["tolerance"]
    e_dense, t1_dense = dense_lccsd(inp)
    assert abs(e_dense - g_corr) < tol
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(Program(lw.PROGRAM_LCCSD), be, inp["segs"], index_base=inp["index_base"])
    _, hist = lw.converge(w, be.value, max_iter=120)
    e_corr = hist[-1]
    assert abs(e_corr - g_corr) < tol                      # lccsd_correlation -0.12865706498547 @ 1e-10
    assert abs(e_corr + inp["e_scf"] - g_total) < tol      # lccsd_energy -75.71298380772593 @ 1e-10
    assert abs(e_corr - e_dense) < 1e-11
    t1 = qm.join_blocks(inp["arrays"]["t1a_old"], [inp["segs"]["v"], inp["segs"]["o"]])
    assert np.max(np.abs(t1 - t1_dense)) < 1e-9 and np.max(np.abs(t1_dense)) > 1e-4    # the singles really are non-zero
    assert not w.locals                                    # every allocated local array was deallocated


# ---------------------------------------------------------------------------------------------------------------------
# CCSD (tests/golden/ccsd_program.sialx = src/sialx/qm/cc/rccsd_rhf.sialx): BASELINE config 3's ground-state energy
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["all_dat", "all_fine"])
def test_ccsd_energy_on_the_oracle_backend_matches_the_reference_golden(oracle, case):
    """the reference's CCSD program (22 procedures: tau, Fae / Fmi / Fme, T1, Wminj, the t1-dressed AO ladder, Wmebj /
    Wmjbe ring terms, ...) against ccsd_energy of eom_test (-75.71251002928709, cc_conv 1e-12; test_qm.cpp:252-253) and
    of eom_ccsd_water_test (-75.71251002936883: the same run stopped at cc_conv 1e-10; :990-991)"""
    inp = lw.inputs(case)
    g_tight, g_loose = lw.golden_ccsd()
    tol = lw.GOLDEN["tolerance"]
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(Program(lw.PROGRAM_CCSD), be, inp["segs"], index_base=inp["index_base"])
    e_mp2, hist = lw.converge(w, be.value, max_iter=150)
    e_total = hist[-1] + inp["e_scf"]
    assert abs(e_mp2 + inp["e_scf"] - lw.GOLDEN["all_electron"]["mp2_energy"]) < tol
    assert abs(e_total - g_tight) < 1e-11          # measured: 6.3e-13
    assert abs(e_total - g_loose) < tol            # the reference's own 1e-10 around its loosely converged value
    t1 = qm.join_blocks(inp["arrays"]["t1a_old"], [inp["segs"]["v"], inp["segs"]["o"]])
    assert np.max(np.abs(t1)) > 1e-3 and not w.locals


@pytest.mark.parametrize("case", ["hf_dat", "hf_fc_dat", "hf_fc_fine"])
def test_ccsd_energy_of_hydrogen_fluoride_matches_the_reference_golden(oracle, case):
    """a second molecule: the CCSD stage of the reference's enabled CCSD(T) test (second_ccsdpt_test: ccsd_correlation
    -0.12588695910754, ccsd_energy -99.58563872286452, cc_conv 1e-10; test_qm.cpp:86-102) and of lamccsdpt_test (frozen
    core, ccsd_energy -99.583972376431, cc_conv 1e-12; :810-824)"""
    inp = lw.inputs(case)
    g, tol = lw.GOLDEN["hf"], lw.GOLDEN["tolerance"]
    assert abs(inp["e_scf"] - g["scf_energy"]) < tol          # -99.45975176375698: pins the inputs of this molecule
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(Program(lw.PROGRAM_CCSD), be, inp["segs"], index_base=inp["index_base"])
    _, hist = lw.converge(w, be.value, max_iter=150)
    if case == "hf_dat":
        assert abs(hist[-1] - g["ccsd_correlation"]) < tol                       # measured 9.1e-12
        assert abs(hist[-1] + inp["e_scf"] - g["ccsd_energy"]) < tol
    else:
        assert abs(hist[-1] + inp["e_scf"] - g["frozen_core_ccsd_energy"]) < tol   # measured 1.7e-12
    assert not w.locals


# ---------------------------------------------------------------------------------------------------------------------
# several workers: the pardo work distribution, barriers, many-writer accumulates and the collective of the real programs
# ---------------------------------------------------------------------------------------------------------------------
class SharedOracleBackend(OracleBackend):
    """OracleBackend for `world` walkers running as threads on ONE set of arrays: a barrier is a real barrier, put /
    put += are atomic at the owner (the reference serialises them in the server loop, sip_server.cpp:172-334), the
    collective is an all-reduce.  Blocks fetched with get/request are private copies until the next barrier, as in the
    worker-side cache (sial_ops_parallel.cpp:41-47)."""

    def __init__(self, oracle, arrays, shared, **kw):
        super().__init__(oracle, arrays, **kw)
        self.sh, self.cache = shared, {}

    def array_block(self, name, segs, shape):
        key = (name, segs)
        if key not in self.cache:
            with self.sh["lock"]:
                self.cache[key] = HostBlockCopy(super().array_block(name, segs, shape).a)
        return self.cache[key]

    def put(self, arr, segs, b):
        with self.sh["lock"]:
            super().put(arr, segs, b)

    def put_accumulate(self, arr, segs, b):
        with self.sh["lock"]:
            super().put_accumulate(arr, segs, b)

    def put_initialize(self, arr, segs, shape, v):
        with self.sh["lock"]:
            super().put_initialize(arr, segs, shape, v)

    def barrier(self):
        self.sh["barrier"].wait()
        self.cache.clear()

    def collective_sum(self, a, b):
        with self.sh["lock"]:
            self.sh["sum"].append(b)
        self.sh["barrier"].wait()
        total = sum(self.sh["sum"])
        self.sh["barrier"].wait()
        with self.sh["lock"]:
            self.sh["sum"].clear()
        self.sh["barrier"].wait()
        return a + total


def HostBlockCopy(a):
    from sial_oracle_backend import HostBlock
    return HostBlock(np.array(a, order="F"))


@pytest.mark.parametrize("program,world", [("lccd", 3), ("lccsd", 2), ("ccsd", 3)])
def test_programs_on_several_workers_reproduce_the_goldens(oracle, program, world):
    """`world` walkers (threads) share the arrays; every walker runs the whole program and executes the pardo
    iterations k with (k - 1) mod world == rank (loop_manager.cpp:468-499).  The converged energy must be the
    single-worker one -- which it only is if every read of an array another worker writes is separated from the
    write by a barrier in the program text, i.e. if the transcribed programs kept the reference's barrier structure."""
    import threading

    text = {"lccd": lw.PROGRAM, "lccsd": lw.PROGRAM_LCCSD, "ccsd": lw.PROGRAM_CCSD}[program]
    inp = lw.inputs("all_fine")
    g = lw.GOLDEN["all_electron"]
    want = {"lccd": g["lccd_energy"], "lccsd": g["lccsd_energy"], "ccsd": g["ccsd_energy"]}[program]
    shared = {"lock": threading.Lock(), "barrier": threading.Barrier(world), "sum": []}
    prog = Program(text)
    out, errs = [None] * world, []

    def run(rank):
        try:
            be = SharedOracleBackend(oracle, inp["arrays"], shared, fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
            w = Walker(prog, be, inp["segs"], rank=rank, world=world, index_base=inp["index_base"])
            w.run()
            e = None
            for _ in range(12):                      # a fixed number of iterations on every worker: no divergence
                e = be.value(w.run_proc("iteration")["ecorrab"])
            out[rank] = (e, be.calls)
        except BaseException as ex:                  # noqa: BLE001 -- release the others, then report
            errs.append(ex)
            shared["barrier"].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs
    # single worker, same number of iterations
    inp1 = lw.inputs("all_fine")
    be1 = OracleBackend(oracle, inp1["arrays"], fock=inp1["fock"], moa_seg_ranges=inp1["moa_seg_ranges"])
    w1 = Walker(prog, be1, inp1["segs"], index_base=inp1["index_base"])
    w1.run()
    for _ in range(12):
        e1 = be1.value(w1.run_proc("iteration")["ecorrab"])
    assert all(abs(o[0] - e1) < 1e-12 for o in out), (out, e1)
    assert abs(e1 + inp1["e_scf"] - want) < 1e-5          # 12 plain iterations: converging to the golden
    calls = [o[1] for o in out]
    assert min(calls) > 0.5 * max(calls)                   # the work really is shared


@pytest.mark.parametrize("case", ["hf_dat", "hf_fine"])
def test_ccsd_t_energy_of_hydrogen_fluoride_matches_the_reference_golden(oracle, case):
    """the reference's enabled CCSD(T) test (second_ccsdpt_test, test_qm.cpp:86-126): ccsdpt_energy -99.58619978246637.
    CCSD by the reference's program (ccsd_program.sialx); the (T) correction by a closed-shell restatement in the same
    statement subset (ccsd_t_restated.sialx -- NOT the reference's triples programs, see its header) whose rank-6
    contractions, outer products, permutes, denominator and scalar contractions all go through the block backend"""
    inp = lw.inputs(case)
    g, tol = lw.GOLDEN["hf"], lw.GOLDEN["tolerance"]
    assert abs(g["ccsd_energy"] + g["eaaa"] + g["esaaa"] + g["eaab"] + g["esaab"] - g["ccsdpt_energy"]) < 1e-13
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(Program(lw.PROGRAM_CCSD), be, inp["segs"], index_base=inp["index_base"])
    _, hist = lw.converge(w, be.value, max_iter=150)
    sc = Walker(Program(lw.PROGRAM_PT), be, inp["segs"], index_base=inp["index_base"]).run()
    e_t = be.value(sc["et"])
    assert abs(e_t - (g["ccsdpt_energy"] - g["ccsd_energy"])) < 1e-11          # E(T) itself: measured 1.3e-13
    for name in ("eaaa", "esaaa", "eaab", "esaab"):                             # the four numbers the reference asserts:
        assert abs(be.value(sc[name]) - g[name]) < 1e-11, name                  # measured 6e-15, 7e-15, 2.6e-13, 1.1e-13
    assert abs(hist[-1] + inp["e_scf"] + e_t - g["ccsdpt_energy"]) < tol       # measured 9.5e-12 (reference cc_conv 1e-10)


def test_d_shell_integrals_reproduce_the_neon_scf_energy():
    """ccsdpt_test.dat (BASELINE config 2) is a neon atom in a [3s2p1d] set with SPHERICAL d functions -- the only setup of
    the energy tests with l = 2.  Pins of the input stage for it: 14 functions (3 + 2*3 + 5), the RHF energy of Ne / cc-pVDZ
    (-128.48877555 in the basis-set literature; the setup's exponents are that set's), a five-fold degenerate d level and
    three-fold degenerate p levels (what a wrong solid-harmonic combination or cartesian norm would split)."""
    setup, basis, S, eri, e_nuc, e_scf, eps, C = lw.scf("ccsdpt_test.dat")
    assert S.shape == (14, 14) and e_nuc == 0
    assert abs(e_scf - (-128.48877555)) < 2e-8
    assert np.ptp(eps[2:5]) < 1e-9 and np.ptp(eps[5:8]) < 1e-9 and np.ptp(eps[9:14]) < 1e-9


@pytest.mark.parametrize("case", ["ne_dat", "ne_fine"])
def test_ccsd_t_of_neon_matches_the_goldens_of_ccsdpt_test(oracle, case):
    """BASELINE config 2 at file level: test/ccsdpt_test.dat (DISABLED_ccsdpt_test, test/test_qm.cpp:22-52) asserts
    eaab -0.0010909774775509193 and esaab 8.5547845910409156e-05 after scf / tran / rccsd / rccsdpt_aaa / rccsdpt_aab.
    The setup stops its CCSD at cc_conv 1e-7 (and the SCF at 1e-8), so its goldens carry that run's convergence error; the
    tightly converged run here lands 1.5e-10 (eaab) and 2.2e-10 (esaab) from them -- inside north_star's 1e-9 Hartree,
    asserted at 1e-9.  CCSD by the reference's program, (T) by the closed-shell restatement (see the test above), every block
    operation through the backend; `ne_fine` cuts the same orbitals into 2+3 occupied / 4+5 virtual / 3+6+5 AO segments."""
    inp = lw.inputs(case)
    g = lw.GOLDEN["ne_ccsdpt_test"]
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    w = Walker(Program(lw.PROGRAM_CCSD), be, inp["segs"], index_base=inp["index_base"])
    _, hist = lw.converge(w, be.value, max_iter=150)
    sc = Walker(Program(lw.PROGRAM_PT), be, inp["segs"], index_base=inp["index_base"]).run()
    assert abs(be.value(sc["eaab"]) - g["eaab"]) < 1e-9                          # measured -1.5e-10
    assert abs(be.value(sc["esaab"]) - g["esaab"]) < 1e-9                        # measured +2.2e-10
    # segmentation independence to rounding: both cases give the same numbers (measured: 2e-19 apart)
    assert abs(be.value(sc["eaab"]) - (-0.0010909776279972)) < 1e-13
    assert abs(hist[-1] - (-0.190861375509551)) < 1e-11                          # CCSD correlation energy of this run
    assert abs(be.value(sc["et"]) - sum(be.value(sc[k]) for k in ("eaaa", "esaaa", "eaab", "esaab"))) < 1e-15


# ---------------------------------------------------------------------------------------------------------------------
# the integral transformation (tests/golden/tran_program.sialx = src/sialx/qm/utility/tran_rhf_no4v.sialx) in front of it
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,program", [("fine", "lccd"), ("all_fine", "ccsd")])
def test_transformation_program_then_cc_program_reproduce_the_goldens(oracle, case, program):
    """AO integrals + MO coefficients -> [the reference's transformation program] -> MO classes -> [the reference's CC
    program] -> golden energy, every block operation of both programs on the backend.  The classes must equal the
    dense numpy transformation of oracle/qm_inputs.py (which the other tests feed in directly)."""
    inp = lw.inputs(case)
    want = {n: inp["arrays"][n] for n in lw.MO_CLASSES}
    for n in lw.MO_CLASSES:
        inp["arrays"][n] = {}
    be = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    wt = Walker(Program(lw.PROGRAM_TRAN), be, inp["segs"], index_base=inp["index_base"])
    wt.run()
    assert not wt.locals
    for n in lw.MO_CLASSES:
        assert set(inp["arrays"][n]) == set(want[n]), n
        for idx, b in want[n].items():
            assert np.max(np.abs(inp["arrays"][n][idx] - b)) < 1e-13, (n, idx)
    text = {"lccd": lw.PROGRAM, "ccsd": lw.PROGRAM_CCSD}[program]
    _, hist = lw.converge(Walker(Program(text), be, inp["segs"], index_base=inp["index_base"]), be.value, max_iter=150)
    g = lw.golden(case)[1] if program == "lccd" else lw.golden_ccsd()[0]
    assert abs(hist[-1] + inp["e_scf"] - g) < lw.GOLDEN["tolerance"]


def test_transformation_program_on_several_workers(oracle):
    """the transformation program shared by three workers (threads, real barriers, atomic prepare +=): same classes"""
    import threading

    world = 3
    inp = lw.inputs("all_fine")
    want = {n: inp["arrays"][n] for n in lw.MO_CLASSES}
    for n in lw.MO_CLASSES:
        inp["arrays"][n] = {}
    shared = {"lock": threading.Lock(), "barrier": threading.Barrier(world), "sum": []}
    prog, errs, calls = Program(lw.PROGRAM_TRAN), [], [0] * world

    def run(rank):
        try:
            be = SharedOracleBackend(oracle, inp["arrays"], shared, fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
            Walker(prog, be, inp["segs"], rank=rank, world=world, index_base=inp["index_base"]).run()
            calls[rank] = be.calls
        except BaseException as ex:                  # noqa: BLE001
            errs.append(ex)
            shared["barrier"].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for n in lw.MO_CLASSES:
        for idx, b in want[n].items():
            assert np.max(np.abs(inp["arrays"][n][idx] - b)) < 1e-13, (n, idx)
    assert min(calls) > 0.5 * max(calls)


def test_programs_chained_through_persistent_arrays(oracle):
    """transformation program -> `set_persistent` -> a NEW backend with freshly declared arrays -> `restore_persistent` in
    the CC program's READ_2EL -> LCCD golden: the hand-over the reference uses between its programs"""
    tran_text, cc_text = lw.chained_through_persistence(lw.PROGRAM_TRAN, lw.PROGRAM)
    inp = lw.inputs("fine")
    for n in lw.MO_CLASSES:
        inp["arrays"][n] = {}
    OracleBackend.registry.clear()
    be1 = OracleBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    Walker(Program(tran_text), be1, inp["segs"], index_base=inp["index_base"]).run()
    assert sorted(OracleBackend.registry) == sorted(lw.PERSISTED)
    assert not any(x.lower() in inp["arrays"] for x in lw.PERSISTED)        # handed over, not copied
    arrays2 = {n: {} for n in lw.KINDS}                                      # the next program's own declarations
    arrays2["ca"], arrays2["aoint"] = inp["arrays"]["ca"], inp["arrays"]["aoint"]
    be2 = OracleBackend(oracle, arrays2, fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"])
    _, hist = lw.converge(Walker(Program(cc_text), be2, inp["segs"], index_base=inp["index_base"]), be2.value)
    assert sorted(OracleBackend.registry) == ["VSpipi", "Vaaai"]            # persisted for other programs (LCCSD, CCSD)
    assert abs(hist[-1] - lw.GOLDEN["lccd_correlation"]) < lw.GOLDEN["tolerance"]
    with pytest.raises(KeyError):                                            # nothing left to restore a second time
        Walker(Program(cc_text), be2, inp["segs"], index_base=inp["index_base"]).run()


# ---------------------------------------------------------------------------------------------------------------------
# the reference's race rules (distributed_block_consistency.cpp:25-175) applied to the transcribed programs
# ---------------------------------------------------------------------------------------------------------------------
class AccessLoggingBackend(OracleBackend):
    """logs (section, worker, GET/PUT/PUT_ACCUMULATE) of every block of a served / distributed array"""
    GET, PUT, ACC = 0, 1, 2

    def __init__(self, *a, log, rank, **kw):
        super().__init__(*a, **kw)
        self.log, self.rank, self.section = log, rank, 1

    def array_block(self, name, segs, shape):
        self.log[name, segs].append((self.section, self.rank, self.GET))
        return super().array_block(name, segs, shape)

    def static_block(self, name, segs, shape):          # static arrays are replicated, not served: no rule applies
        return OracleBackend.array_block(self, name, segs, shape)

    def put(self, arr, segs, b):
        self.log[arr, segs].append((self.section, self.rank, self.PUT))
        super().put(arr, segs, b)

    def put_accumulate(self, arr, segs, b):
        self.log[arr, segs].append((self.section, self.rank, self.ACC))
        super().put_accumulate(arr, segs, b)

    def put_initialize(self, arr, segs, shape, v):
        self.log[arr, segs].append((self.section, self.rank, self.PUT))
        super().put_initialize(arr, segs, shape, v)

    def barrier(self):
        self.section += 1


def illegal_accesses(oracle, text, case, world=3):
    import collections

    log = collections.defaultdict(list)
    for r in range(world):          # the access pattern does not depend on the data: the workers can run one after another
        inp = lw.inputs(case)
        be = AccessLoggingBackend(oracle, inp["arrays"], fock=inp["fock"], moa_seg_ranges=inp["moa_seg_ranges"], log=log, rank=r)
        w = Walker(Program(text), be, inp["segs"], rank=r, world=world, index_base=inp["index_base"])
        w.run()
        if "iteration" in w.p.procs:
            w.run_proc("iteration")
    bad = []
    for key, acc in log.items():
        acc.sort(key=lambda t: (t[0], t[1]))
        k = oracle.block_consistency([a[2] for a in acc], [a[1] for a in acc], [a[0] for a in acc])
        if k != -1:
            bad.append((key, acc[k]))
    return bad, sum(len(v) for v in log.values())


@pytest.mark.parametrize("program,case", [("tran", "all_fine"), ("lccd", "fine"), ("lccsd", "all_fine"), ("ccsd", "all_fine"),
                                          ("(T)", "hf_fine")])
def test_transcribed_programs_obey_the_reference_race_rules(oracle, program, case):
    """every access of three workers to every block of every served / distributed array, section by section, through
    the oracle's restatement of DistributedBlockConsistency::update_and_check_consistency: no illegal access, i.e. the
    transcriptions kept the barrier structure that makes the reference's programs race-free"""
    text = {"tran": lw.PROGRAM_TRAN, "lccd": lw.PROGRAM, "lccsd": lw.PROGRAM_LCCSD, "ccsd": lw.PROGRAM_CCSD,
            "(T)": lw.PROGRAM_PT}[program]
    bad, n = illegal_accesses(oracle, text, case)
    assert n > 2000 and not bad, bad[:5]


def test_race_rule_check_detects_a_missing_barrier(oracle):
    """negative control: without the barrier between the first and second quarter transformation, Vxxxi is written by
    one worker and read by another in the same section"""
    text = lw.PROGRAM_TRAN.replace("endpardo mu, nu, lambda\nserver_barrier\n", "endpardo mu, nu, lambda\n", 1)
    assert text != lw.PROGRAM_TRAN
    bad, _ = illegal_accesses(oracle, text, "all_fine")
    assert bad and all(key[0] == "vxxxi" for key, _ in bad)
