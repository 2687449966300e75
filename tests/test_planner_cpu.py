"""CPU tests of the HOST logic behind the kernels: the Shape (stride groups) the fused contraction kernel consumes
and the tile/offset tables of the permute kernel are replayed here in numpy exactly as the kernels use them
(address = sum of per-group offsets) and compared with the oracle.  No GPU needed."""
import ctypes as C
import itertools
import random

import numpy as np
import pytest


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


def get_shape(sip, ptrn, lext, rext, dext):
    L = sip.lib()
    n = L.sipgpu_debug_shape_ints()
    buf = (C.c_int * n)()
    rc = L.sipgpu_debug_contract_shape(sip._ia(ptrn), len(lext), sip._ia(lext), len(rext), sip._ia(rext), len(dext),
                                       sip._ia(dext), buf)
    v = list(buf)
    if rc != 0:
        return rc, None
    sh = dict(M=v[0], N=v[1], K=v[2], nm=v[3], nn=v[4], nk=v[5])
    names = ["mext", "msL", "msD", "next", "nsR", "nsD", "kext", "ksL", "ksR"]
    for i, nme in enumerate(names):
        sh[nme] = v[6 + 6 * i: 12 + 6 * i]
    sh["a_kc"], sh["b_kc"] = v[60], v[61]
    return 0, sh


def offsets(n, nd, ext, s0, s1):
    """decompose2 of contract.cu for every linear index 0..n-1"""
    lin = np.arange(n)
    o0 = np.zeros(n, dtype=np.int64)
    o1 = np.zeros(n, dtype=np.int64)
    for i in range(nd):
        r = lin % ext[i]
        lin = lin // ext[i]
        o0 += r * s0[i]
        o1 += r * s1[i]
    return o0, o1


def replay_contract(sh, L, R, dsize):
    Lf, Rf = L.ravel(order="F"), R.ravel(order="F")
    mL, mD = offsets(sh["M"], sh["nm"], sh["mext"], sh["msL"], sh["msD"])
    nR, nD = offsets(sh["N"], sh["nn"], sh["next"], sh["nsR"], sh["nsD"])
    kL, kR = offsets(sh["K"], sh["nk"], sh["kext"], sh["ksL"], sh["ksR"])
    A = Lf[kL[:, None] + mL[None, :]]  # K x M
    B = Rf[kR[:, None] + nR[None, :]]  # K x N
    Dp = A.T @ B
    D = np.full(dsize, np.nan)
    D[(mD[:, None] + nD[None, :]).ravel()] = Dp.ravel()
    return D


def random_pattern(pyrng, exts):
    nfl, nfr, nc = pyrng.randint(1, 3), pyrng.randint(1, 3), pyrng.randint(0, 3)
    labels = list(range(1, nfl + nfr + nc + 1))
    ext = {lab: pyrng.choice(exts) for lab in labels}
    fl, fr, cc = labels[:nfl], labels[nfl:nfl + nfr], labels[nfl + nfr:]
    llab, rlab, dlab = fl + cc, fr + cc, fl + fr
    pyrng.shuffle(llab), pyrng.shuffle(rlab), pyrng.shuffle(dlab)
    return dlab, llab, rlab, ext


def test_contract_shape_replay_matches_oracle(sip, oracle):
    pyrng = random.Random(42)
    rng = np.random.default_rng(42)
    for trial in range(300):
        dlab, llab, rlab, ext = random_pattern(pyrng, (1, 2, 3, 4, 5, 7))
        lext, rext, dext = [ext[x] for x in llab], [ext[x] for x in rlab], [ext[x] for x in dlab]
        ptrn, ierr = sip.get_contraction_ptrn(dlab, llab, rlab)
        assert ierr == 0
        rc, sh = get_shape(sip, ptrn, lext, rext, dext)
        assert rc == 0
        Lb = np.asfortranarray(rng.uniform(-1, 1, size=lext))
        Rb = np.asfortranarray(rng.uniform(-1, 1, size=rext))
        ref, oerr = oracle.contract_labels(dlab, dext, llab, Lb, rlab, Rb)
        assert oerr == 0
        got = replay_contract(sh, Lb, Rb, int(np.prod(dext)))
        assert not np.any(np.isnan(got)), "every destination element must be written exactly once"
        assert np.allclose(got, ref.ravel(order="F"), rtol=0, atol=1e-12 * max(1.0, np.max(np.abs(ref))))
        # dims and flags agree with determine_index_permutations (F90:799-859)
        _, _, _, dims, _ = oracle.determine_index_permutations(ptrn, lext, rext, dext)
        assert [sh["M"], sh["N"], sh["K"]] == dims
        # stride-1 flags: K-contiguous means the first contracted dim has unit stride in that operand
        if sh["a_kc"]:
            assert sh["ksL"][0] == 1
        if sh["b_kc"]:
            assert sh["ksR"][0] == 1


def test_contract_shape_collapses_gemm_like_patterns(sip):
    # Z[a,i,b,j] = T[a,i,c,k] * V[c,k,b,j] is a plain GEMM: every group collapses to ONE dimension
    v, o = 50, 20
    ptrn, _ = sip.get_contraction_ptrn([1, 2, 3, 4], [1, 2, 5, 6], [5, 6, 3, 4])
    rc, sh = get_shape(sip, ptrn, [v, o, v, o], [v, o, v, o], [v, o, v, o])
    assert rc == 0 and (sh["nm"], sh["nn"], sh["nk"]) == (1, 1, 1)
    assert (sh["M"], sh["N"], sh["K"]) == (1000, 1000, 1000)
    assert sh["msL"][0] == 1 and sh["ksL"][0] == 1000 and sh["ksR"][0] == 1 and sh["nsR"][0] == 1000
    assert sh["a_kc"] == 0 and sh["b_kc"] == 1


def test_contract_shape_rejects_bad_extents(sip):
    rc, _ = get_shape(sip, [1, -1, -2, 2], [3, 4], [5, 6], [3, 6])
    assert rc == 1  # contr_ptrn_ok (F90:861-896) -> ierr 1


def replay_permute(sip, a, transp):
    L = sip.lib()
    rank = a.ndim
    cap = 4096
    meta = (C.c_longlong * 32)()
    rt = (C.c_int * (2 * cap))()
    wt = (C.c_int * (3 * cap))()
    rc = L.sipgpu_debug_permute_plan(rank, sip._ia(a.shape), sip._ia(transp), meta, rt, wt, cap)
    assert rc == 0
    r, V, ntiles = meta[0], meta[1], meta[2]
    src = a.ravel(order="F")
    if r <= 1:
        return src.copy(), 0
    rag = [(meta[3 + 3 * i], meta[4 + 3 * i], meta[5 + 3 * i]) for i in range(2)]
    ntile, tin, tout = list(meta[9:15]), list(meta[15:21]), list(meta[21:27])
    rt = np.array(rt[: 2 * V]).reshape(V, 2)
    wt = np.array(wt[: 3 * V]).reshape(V, 3)
    out = np.full(src.size, np.nan)
    for tile in range(ntiles):
        t, bin_, bout, lim = tile, 0, 0, [1 << 16, 1 << 16]
        for d in range(r):
            c = t % ntile[d]
            t //= ntile[d]
            bin_ += c * tin[d]
            bout += c * tout[d]
            for q in range(2):
                if d == rag[q][0]:
                    lim[q] = min(rag[q][1], rag[q][2] - c * rag[q][1])
        sm = np.full(V, np.nan)
        okr = ((rt[:, 1] & 0xFFFF) < lim[0]) & ((rt[:, 1] >> 16) < lim[1])
        sm[np.nonzero(okr)[0]] = src[bin_ + rt[okr, 0]]
        okw = ((wt[:, 2] & 0xFFFF) < lim[0]) & ((wt[:, 2] >> 16) < lim[1])
        vals = sm[wt[okw, 1]]
        assert not np.any(np.isnan(vals))
        assert np.all(np.isnan(out[bout + wt[okw, 0]])), "an output element written twice"
        out[bout + wt[okw, 0]] = vals
    return out, V


@pytest.mark.parametrize("shape", [(16, 16, 16, 16), (13, 30, 50, 7), (5, 8, 9, 5), (50, 20, 50, 20), (33, 3, 1, 65)])
def test_permute_plan_replay_rank4(sip, oracle, shape):
    rng = np.random.default_rng(1)
    a = np.asfortranarray(rng.uniform(-1, 1, size=shape))
    for perm in itertools.permutations(range(4)):
        transp = [1] + [p + 1 for p in perm]
        got, V = replay_permute(sip, a, transp)
        assert np.array_equal(got, oracle.block_copy(a, transp).ravel(order="F")), (shape, perm)
        assert V <= 4096


def test_permute_plan_replay_other_ranks(sip, oracle):
    pyrng = random.Random(9)
    rng = np.random.default_rng(9)
    for trial in range(80):
        rank = pyrng.randint(1, 6)
        shape = tuple(pyrng.choice([1, 2, 3, 5, 7, 11, 16, 40]) for _ in range(rank))
        if np.prod(shape) > 400000:
            continue
        perm = list(range(rank))
        pyrng.shuffle(perm)
        a = np.asfortranarray(rng.uniform(-1, 1, size=shape))
        transp = [1] + [p + 1 for p in perm]
        got, _ = replay_permute(sip, a, transp)
        assert np.array_equal(got, oracle.block_copy(a, transp).ravel(order="F")), (shape, perm)
