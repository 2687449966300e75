"""CPU tests of the SIAL front-end (aces4_b200/sial_frontend.py): the parser, the walker's loop/scope/where
semantics and the pardo work distribution, executed on the ORACLE backend and compared with dense numpy.einsum of the
LCCD doubles equations.  No GPU needed.  (The same SIAL text runs on the device backend in tests/test_gpu_worklist.py.)"""
import os

import numpy as np
import pytest

from aces4_b200.sial_frontend import Program, SialSyntaxError, Walker
from sial_oracle_backend import OracleBackend

HERE = os.path.dirname(os.path.abspath(__file__))
LCCD = open(os.path.join(HERE, "golden", "lccd_doubles.sialx")).read()
KINDS = {"vpiqj": "vovo", "voooo": "oooo", "viaai": "ovvo", "vaaii": "vvoo", "t2old_ab": "vovo", "t2new_ab": "vovo"}
TAGS = {"vpiqj": 2, "voooo": 3, "viaai": 5, "vaaii": 6, "t2old_ab": 1}


def make_arrays(oracle, segs, seed=0xACE54):
    """name -> {segment tuple: block}, seeded like the synthetic workload (oracle.fill_hash)"""
    arrays = {}
    for name, kind in KINDS.items():
        blocks = {}
        nseg = [len(segs[k]) for k in kind]
        for idx in np.ndindex(*nseg):
            idx1 = tuple(i + 1 for i in idx)
            shape = tuple(segs[k][i] for k, i in zip(kind, idx))
            if name == "t2new_ab":
                blocks[idx1] = np.zeros(shape, order="F")
            else:
                number = 0
                for p in range(4):
                    number = number * nseg[p] + idx[p]
                blocks[idx1] = oracle.fill_hash(shape, seed, (TAGS[name] << 40) | number, 0.1)
        arrays[name] = blocks
    return arrays


def dense(blocks, kind, segs):
    offs = []
    for k in kind:
        o = [0]
        for s in segs[k]:
            o.append(o[-1] + s)
        offs.append(o)
    full = np.zeros([o[-1] for o in offs])
    for idx, b in blocks.items():
        sl = tuple(slice(offs[d][idx[d] - 1], offs[d][idx[d]]) for d in range(len(kind)))
        full[sl] = b
    return full


def dense_reference(arrays, segs):
    V = dense(arrays["vpiqj"], "vovo", segs)
    Vo = dense(arrays["voooo"], "oooo", segs)
    Via = dense(arrays["viaai"], "ovvo", segs)
    Vaa = dense(arrays["vaaii"], "vvoo", segs)
    T = dense(arrays["t2old_ab"], "vovo", segs)
    sym = lambda X: X + np.transpose(X, (2, 3, 0, 1))  # noqa: E731
    new = sym(0.5 * V)
    new += np.einsum("akbl,ikjl->aibj", T, Vo)
    TY = np.einsum("iack->aick", Via) - np.einsum("caik->aick", Vaa)
    new += sym(np.einsum("aick,ckbj->aibj", TY, T))
    W = np.einsum("ckai->ckia", T) - np.einsum("ciak->ckia", T)
    new += sym(np.einsum("ckia,iabj->ckbj", W, Via))
    new += sym(-np.einsum("akcj,bcki->aibj", T, Vaa))
    e = np.einsum("aibj,aibj->", new, 2.0 * V - np.transpose(V, (0, 3, 2, 1)))
    return new, e


def test_parser_reads_the_lccd_fragment():
    p = Program(LCCD)
    assert p.index_kind["a1"] == "v" and p.index_kind["j1"] == "o"
    assert p.arrays["t2new_ab"] == ("served", ("a", "i", "b", "j"))
    pardos = [s for s in p.body if s[0] == "pardo"]
    assert [s[1] for s in pardos] == [("a", "b", "i", "j"), ("a", "b", "i1", "j1"), ("j", "b", "a", "i"),
                                      ("i1", "a1", "a", "i"), ("a", "j", "i1", "b1"), ("a", "i", "b", "j")]
    hh = pardos[1][2]
    assert hh[0] == ("request", "t2old_ab", ("a", "i1", "b", "j1"))
    inner = hh[1][2][0][2]      # do i / do j body
    assert inner[1] == ("contract", "taibj", ("a", "i", "b", "j"), "t2old_ab", ("a", "i1", "b", "j1"), "voooo",
                        ("i", "i1", "j", "j1"))
    assert inner[2] == ("put", "t2new_ab", ("a", "i", "b", "j"), "+=", "taibj", ("a", "i", "b", "j"))


@pytest.mark.parametrize("bad", ["pardo a\n  Taibj[a] = 1.0\n", "x[i] = y[i] / z[i]\n", "where a ** b\n",
                                 "moaindex q = 1 7\n", "enddo i\n", "prepare A[i] -= T[i]\n"])
def test_parser_rejects_what_it_does_not_understand(bad):
    with pytest.raises(SialSyntaxError):
        Program(bad)


@pytest.mark.parametrize("segs", [{"o": [2, 3], "v": [3, 4]}, {"o": [3], "v": [2, 2, 3]}])
def test_lccd_doubles_on_the_oracle_backend_match_dense_einsum(oracle, segs):
    arrays = make_arrays(oracle, segs)
    want, e_want = dense_reference(arrays, segs)
    be = OracleBackend(oracle, arrays)
    scal = Walker(Program(LCCD), be, segs).run()
    got = dense(arrays["t2new_ab"], "vovo", segs)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    assert abs(be.value(scal["ecorrab"]) - e_want) <= 1e-12 * abs(e_want)


def test_pardo_work_distribution_partitions_the_iterations(oracle):
    """two workers, each walking the whole program but executing only its iterations (loop_manager.cpp:468-499):
    the union of their puts is the single-worker result, and each worker ran about half of the block ops"""
    segs = {"o": [2, 3], "v": [3, 4]}
    arrays1 = make_arrays(oracle, segs)
    Walker(Program(LCCD), OracleBackend(oracle, arrays1), segs).run()
    arrays2 = make_arrays(oracle, segs)
    calls, e = [], 0.0
    for r in range(2):
        be = OracleBackend(oracle, arrays2)
        # the energy pardo reads T2new after the barrier: run the residual pardos on both ranks first
        prog = Program(LCCD.split("proc energy")[0])
        Walker(prog, be, segs, rank=r, world=2).run()
        calls.append(be.calls)
    for idx, b in arrays1["t2new_ab"].items():
        assert np.max(np.abs(arrays2["t2new_ab"][idx] - b)) <= 1e-12
    assert abs(calls[0] - calls[1]) <= 0.1 * max(calls) and min(calls) > 0


def test_where_clause_and_iteration_counter(oracle):
    text = """
    moaindex i = baocc: eaocc
    moaindex j = baocc: eaocc
    served A[i,j]
    temp T[i,j]
    pardo i, j
        where i <= j
        T[i,j] = 1.0
        put A[i,j] += T[i,j]
    endpardo i, j
    """
    segs = {"o": [2, 2, 3], "v": [1]}
    seen = []
    for r in range(3):
        arrays = {"a": {}}
        Walker(Program(text), OracleBackend(oracle, arrays), segs, rank=r, world=3).run()
        seen.append(sorted(arrays["a"]))
    every = sorted(x for s in seen for x in s)
    assert every == sorted((i, j) for i in (1, 2, 3) for j in (1, 2, 3) if i <= j)
    # first index fastest, k-th where-true iteration -> worker k mod 3
    order = [(i, j) for j in (1, 2, 3) for i in (1, 2, 3) if i <= j]
    for r in range(3):
        assert seen[r] == sorted(order[r::3])


# ---------------------------------------------------------------------------------------------------------------
# CCSD-shaped statements: outer products, put = number, put += into a rank-2 array, block *= scalar variable
# ---------------------------------------------------------------------------------------------------------------
CCSD = open(os.path.join(HERE, "golden", "ccsd_tau_singles.sialx")).read()
CCSD_KINDS = {"t2old_ab": "vovo", "tau_ab": "vovo", "vpiqj": "vovo", "t1a_old": "vo", "t1a_new": "vo", "fme_a": "vo"}


def make_ccsd_arrays(oracle, segs, seed=0xACE54):
    arrays = {}
    for tag, (name, kind) in enumerate(CCSD_KINDS.items(), 1):
        blocks = {}
        nseg = [len(segs[k]) for k in kind]
        for idx in np.ndindex(*nseg):
            shape = tuple(segs[k][i] for k, i in zip(kind, idx))
            number = 0
            for p in range(len(kind)):
                number = number * nseg[p] + idx[p]
            filled = name in ("t2old_ab", "vpiqj", "t1a_old", "fme_a")
            blocks[tuple(i + 1 for i in idx)] = (oracle.fill_hash(shape, seed, (tag << 40) | number, 0.1) if filled
                                                 else np.full(shape, np.nan, order="F"))
        arrays[name] = blocks
    return arrays


def test_parser_reads_the_ccsd_fragment():
    p = Program(CCSD)
    kinds = [s[0] for s in p.body]
    assert kinds.count("pardo") == 4 and "sset" in kinds and kinds[-1] == "collective"
    tau = p.body[3][2]                                   # the tau pardo (after `half = 0.5`)
    assert [s[0] for s in tau] == ["request", "request", "request", "contract", "assign", "add", "put"]
    assert tau[3][3:] == ("t1a_old", ("a", "i"), "t1a_old", ("b", "j"))          # `^` parsed as a contraction
    assert p.body[0][2][0] == ("put_init", "t1a_new", ("a", "i"), 0.0)
    singles = p.body[5][2][0][2][0][2]                   # pardo a,i / do b / do j
    assert ("scale_by", "tai", ("a", "i"), "half") in singles


@pytest.mark.parametrize("world", [1, 3])
def test_ccsd_fragment_on_the_oracle_backend_matches_dense_einsum(oracle, world):
    segs = {"o": [3, 2], "v": [4, 3, 2]}
    arrays = make_ccsd_arrays(oracle, segs)
    T2 = dense(arrays["t2old_ab"], "vovo", segs)
    V = dense(arrays["vpiqj"], "vovo", segs)
    t1 = dense(arrays["t1a_old"], "vo", segs)
    F = dense(arrays["fme_a"], "vo", segs)
    prog = Program(CCSD)
    esum = 0.0
    # every "worker" walks the same program on the shared arrays, one barrier section at a time (the sections of this
    # program are separated by barriers, so running the workers one after the other per program is NOT equivalent; the
    # oracle backend is sequential, so emulate the sections by running each pardo for all ranks before the next one)
    walkers = [Walker(prog, OracleBackend(oracle, arrays), segs, rank=r, world=world) for r in range(world)]
    for st in prog.body:
        for w in walkers:
            w._block([st])
    for w in walkers:
        esum += w.be.value(w.scalars["esum"])
    tau = T2 + np.einsum("ai,bj->aibj", t1, t1)
    assert np.allclose(dense(arrays["tau_ab"], "vovo", segs), tau, rtol=1e-13, atol=1e-15)
    assert np.allclose(dense(arrays["t1a_new"], "vo", segs), 0.5 * np.einsum("aibj,bj->ai", tau, F), rtol=1e-12, atol=1e-15)
    e_ref = np.einsum("aibj,aibj->", 2.0 * V - np.transpose(V, (0, 3, 2, 1)), tau)
    assert abs(esum - e_ref) <= 1e-12 * abs(e_ref)


# ---------------------------------------------------------------------------------------------------------------------
# procedures, local arrays, p-index arrays, static arrays, absolute index values (used by tests/golden/lcc*_program.sialx)
# ---------------------------------------------------------------------------------------------------------------------
DECL = """
moaindex i = baocc: eaocc
moaindex j = baocc: eaocc
moaindex a = bavirt: eavirt
moaindex p = baocc: eavirt
moaindex q = baocc: eavirt
aoindex mu = 1: norb
served F[p,q]
static ca[mu,p]
served OUT[a,i]
served S[i,j]
temp T[a,i]
temp U[a,i]
local L[a,i]
scalar x
"""


def test_procedures_run_only_when_called_and_fragments_run_in_order(oracle):
    text = DECL + """
    proc one
    x += 1.0
    endproc one
    proc ten
    x += 10.0
    call one
    endproc ten
    """
    frag = Program(text)                       # procedures only: a fragment, textual order (1 + 10 + 1)
    assert list(frag.procs) == ["one", "ten"]
    assert Walker(frag, OracleBackend(oracle, {}), {"o": [1], "v": [1]}).run()["x"] == 12.0
    main = Program(text + "call ten\ncall ten\n")   # with a main program only the calls run
    w = Walker(main, OracleBackend(oracle, {}), {"o": [1], "v": [1]})
    assert w.run()["x"] == 22.0
    assert w.run_proc("one")["x"] == 23.0
    with pytest.raises(SialSyntaxError):
        w.run_proc("nope")


@pytest.mark.parametrize("bad", ["proc a\nproc b\nendproc b\nendproc a\n", "endproc a\n", "proc a\n", "proc a\nendproc a\nproc a\nendproc a\n",
                                 "pardo i\nproc a\nendproc a\nendpardo i\n", "call\n", "allocate L\n"])
def test_parser_rejects_malformed_procedures(bad):
    with pytest.raises(SialSyntaxError):
        Program("moaindex i = baocc: eaocc\n" + bad)


def test_p_index_arrays_static_arrays_and_local_arrays(oracle):
    """F[p,q] addressed with occupied and virtual labels (segment number of a virtual label shifted by the number of
    occupied segments), ca[mu,a] read as a slice of a static array, a local array that outlives `do` scopes and is
    zero-filled on allocation"""
    segs = {"o": [2, 1], "v": [3, 2], "ao": [4]}
    rng = np.random.default_rng(5)
    n = 8
    Fd = rng.uniform(-1, 1, (n, n))
    cad = rng.uniform(-1, 1, (4, n))
    from oracle import qm_inputs as qm
    pseg = segs["o"] + segs["v"]
    arrays = {"f": qm.split_blocks(Fd, [pseg, pseg]), "ca": qm.split_blocks(cad, [segs["ao"], pseg]), "out": {}, "s": {}}
    text = DECL + """
    pardo a, i
        request F[a,i]
        allocate L[a,*]
        do j
            request F[a,j]
            T[a,j] = F[a,j]
            L[a,j] += T[a,j]
            L[a,j] += T[a,j]
        enddo j
        T[a,i]  = L[a,i]
        U[a,i]  = F[a,i]
        T[a,i] -= U[a,i]
        prepare OUT[a,i] = T[a,i]
        deallocate L[a,*]
    endpardo a, i
    pardo i, j
        request F[i,j]
        prepare S[i,j] = F[i,j]
    endpardo i, j
    """
    w = Walker(Program(text), OracleBackend(oracle, arrays), segs)
    w.run()
    assert not w.locals
    got = qm.join_blocks(arrays["out"], [segs["v"], segs["o"]])
    assert np.max(np.abs(got - Fd[3:, :3])) < 1e-15   # 2F - F
    assert np.array_equal(qm.join_blocks(arrays["s"], [segs["o"], segs["o"]]), Fd[:3, :3])
    # an ao label cannot address a p dimension, and a local array must be declared `local`
    with pytest.raises(SialSyntaxError):
        Walker(Program(DECL + "pardo mu, i\nrequest F[mu,i]\nendpardo mu, i\n"), OracleBackend(oracle, arrays), segs).run()
    with pytest.raises(SialSyntaxError):
        Walker(Program(DECL + "pardo a\nallocate T[a,*]\nendpardo a\n"), OracleBackend(oracle, arrays), segs).run()


def test_super_instructions_receive_absolute_index_values(oracle):
    """moa segments [1 | 2 2 | 3 5] with occ = segments 2..3 and virt = 4..5: energy_denominator_rhf must see 2..5"""
    seen = []

    class Spy(OracleBackend):
        def execute(self, fname, blocks, segs, kinds, bare):
            seen.append((fname, segs[0], tuple(kinds[0]), tuple(bare)))

    text = DECL + "pardo a, i\nT[a,i] = 1.0\nexecute energy_denominator_rhf T[a,i] fock_a\nendpardo a, i\n"
    Walker(Program(text), Spy(oracle, {}), {"o": [2, 2], "v": [3, 5]}, index_base={"o": 1, "v": 3}).run()
    assert sorted(s[1] for s in seen) == [(4, 2), (4, 3), (5, 2), (5, 3)]
    assert seen[0][0] == "energy_denominator_rhf" and seen[0][2] == ("v", "o") and seen[0][3] == ("fock_a",)


def test_persistence_statements(oracle):
    text = DECL + 'x = 2.5\nset_persistent x "ex"\nset_persistent S "Sij"\n'
    arrays = {"s": {(1, 1): np.ones((2, 2), order="F")}, "f": {}, "ca": {}, "out": {}}
    OracleBackend.registry.clear()
    Walker(Program(text), OracleBackend(oracle, arrays), {"o": [2], "v": [1], "ao": [1]}).run()
    assert "s" not in arrays and OracleBackend.registry["ex"] == 2.5
    arrays2 = {"s": {}, "f": {}, "ca": {}, "out": {}}
    w = Walker(Program(DECL + 'restore_persistent S "Sij"\nrestore_persistent x "ex"\n'), OracleBackend(oracle, arrays2),
               {"o": [2], "v": [1], "ao": [1]})
    assert w.run()["x"] == 2.5 and np.array_equal(arrays2["s"][1, 1], np.ones((2, 2)))
    for bad in ('set_persistent S\n', 'set_persistent T "t"\n', 'restore_persistent "x"\n'):
        with pytest.raises(SialSyntaxError):
            Walker(Program(DECL + bad), OracleBackend(oracle, dict(arrays2)), {"o": [2], "v": [1], "ao": [1]}).run()
