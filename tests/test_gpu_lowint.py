"""The bandwidth-shaped contraction kernel (aces4_b200/csrc/lowint.cu) against the oracle's tensor_block_contract_
(tensor_dil_omp.F90:662-796 restated): every class it takes from the 128-wide tile kernel -- rank-2 results of rank-4 blocks
with long contracted ranges (split along K, partial sums through red.add), one segment-sized free + contracted index (M tiled
by 64), matrix-vector shapes (N = 1), tiny matrices, dot products (M = N = 1), chains over contracted segments, ragged extents
that are not multiples of 8 or 4, operands read in place from slices of static arrays -- and the same problems forced through
the tile kernel (sipgpu_set_tuning) so that both kernels stay covered by the small shapes of the rest of the suite."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    yield s.api
    s.api.set_tuning("lowint_scope", 1)


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def rand_block(rng, shape):
    return np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape))


CASES = [
    # (name, dlab, llab, rlab, extents by label)
    ("rank-2 of rank-4, long K", "ab", "aicj", "bicj", dict(a=50, b=50, i=20, c=50, j=20)),
    ("rank-2 of rank-4, k-fast operands", "ab", "cade", "dbce", dict(a=20, b=20, c=24, d=24, e=9)),
    ("rank-2 of rank-4, ragged", "ij", "iakb", "jakb", dict(i=13, j=7, a=11, k=5, b=9)),
    ("skinny, M tiled", "aibj", "aicj", "cb", dict(a=50, i=20, b=50, j=20, c=50)),
    ("skinny, transposing", "abcd", "ecba", "ed", dict(a=9, b=10, c=7, d=13, e=11)),
    ("skinny, small side on L", "aibj", "ca", "cibj", dict(a=20, i=6, b=14, j=5, c=20)),
    ("matrix-vector", "ab", "abcd", "cd", dict(a=50, b=20, c=30, d=8)),
    ("matrix-vector, K-fast", "ab", "cdab", "dc", dict(a=17, b=9, c=12, d=10)),
    ("tiny matrices", "ab", "ac", "cb", dict(a=20, b=20, c=50)),
    ("tiny, odd", "ab", "ca", "bc", dict(a=3, b=5, c=7)),
    ("dot, contiguous", "ab", "cda", "cdb", dict(a=1, b=1, c=50, d=41)),
    ("dot, permuted K", "ab", "cda", "dcb", dict(a=1, b=1, c=6, d=7)),
    ("rank-5 DIIS element", "kl", "aibjk", "aibjl", dict(k=1, l=1, a=12, i=5, b=12, j=5)),
    ("64 x 64", "ab", "ac", "cb", dict(a=64, b=64, c=64)),
    ("65 rows", "ab", "ac", "cb", dict(a=65, b=64, c=33)),
    ("outer product", "ab", "ax", "xb", dict(a=20, b=33, x=1)),
]


@pytest.mark.parametrize("route", ["lowint", "tiles"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_single_blocks(sip, oracle, case, route):
    sip.set_tuning("lowint_scope", 2 if route == "lowint" else 0)
    name, d, l, r, ext = case
    labs = sorted(set(d + l + r))
    num = {c: i + 1 for i, c in enumerate(labs)}
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
    L = rand_block(rng, tuple(ext[c] for c in l))
    R = rand_block(rng, tuple(ext[c] for c in r))
    dext = [ext[c] for c in d]
    ref, ierr = oracle.contract_labels([num[c] for c in d], dext, [num[c] for c in l], L, [num[c] for c in r], R)
    assert ierr == 0
    dL, dR = sip.DeviceBlock.from_numpy(L), sip.DeviceBlock.from_numpy(R)
    got = sip.contract_labels([num[c] for c in d], dext, [num[c] for c in l], dL, [num[c] for c in r], dR).to_numpy()
    assert relerr(got.reshape(ref.shape), ref) <= TOL, name
    # alpha / beta (the fused accumulate of `D += f * L*R`)
    D0 = rand_block(rng, tuple(dext))
    out = sip.DeviceBlock.from_numpy(D0)
    sip.contract_labels([num[c] for c in d], dext, [num[c] for c in l], dL, [num[c] for c in r], dR, out=out, alpha=-0.5, beta=2.0)
    assert relerr(out.to_numpy().reshape(ref.shape), -0.5 * ref + 2.0 * D0) <= TOL, name


@pytest.mark.parametrize("nblocks,chain_max", [(1, 1), (3, 7), (40, 3), (700, 2)])
@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[3], CASES[8], CASES[10]], ids=lambda c: c[0])
def test_work_lists_with_chains(sip, oracle, case, nblocks, chain_max):
    """few destinations (split along the chain and K), many destinations (one item each), chains of unequal length"""
    sip.set_tuning("lowint_scope", 2)
    name, d, l, r, ext = case
    if nblocks >= 40:
        ext = {c: max(1, min(e, 12)) for c, e in ext.items()}   # keep the oracle fast
    labs = sorted(set(d + l + r))
    num = {c: i + 1 for i, c in enumerate(labs)}
    dl_, ll_, rl_ = [num[c] for c in d], [num[c] for c in l], [num[c] for c in r]
    ptrn, ierr = sip.get_contraction_ptrn(dl_, ll_, rl_)
    assert ierr == 0
    rng = np.random.default_rng(nblocks * 31 + chain_max)
    lsh, rsh, dsh = tuple(ext[c] for c in l), tuple(ext[c] for c in r), tuple(ext[c] for c in d)
    npool = min(nblocks * chain_max, 24)
    Lh = [rand_block(rng, lsh) for _ in range(npool)]
    Rh = [rand_block(rng, rsh) for _ in range(npool)]
    Ld, Rd = [sip.DeviceBlock.from_numpy(x) for x in Lh], [sip.DeviceBlock.from_numpy(x) for x in Rh]
    prod = {}

    def pair_product(i, j):
        if (i, j) not in prod:
            prod[(i, j)], e = oracle.contract_labels(dl_, list(dsh), ll_, Lh[i], rl_, Rh[j])
            assert e == 0
        return prod[(i, j)]

    lp, rp, chain, refs, Ds = [], [], [0], [], []
    for b in range(nblocks):
        n = 1 + (b * 7) % chain_max
        ref = np.zeros(dsh, order="F")
        for c in range(n):
            i, j = (b * 3 + c) % npool, (b * 5 + 2 * c + 1) % npool
            lp.append(Ld[i].ptr), rp.append(Rd[j].ptr)
            ref = ref + pair_product(i, j)
        chain.append(len(lp))
        refs.append(ref)
        Ds.append(sip.DeviceBlock(dsh))
    bc = sip.BatchedContraction(ptrn, [lsh] * nblocks, [rsh] * nblocks, [dsh] * nblocks, lp, rp, [x.ptr for x in Ds], chain_start=chain)
    bc.launch()
    step = max(1, nblocks // 25)
    for b in range(0, nblocks, step):
        assert relerr(Ds[b].to_numpy().reshape(refs[b].shape), refs[b]) <= TOL, (name, b)
    bc.launch(alpha=0.25, beta=1.0)
    for b in range(0, nblocks, step):
        assert relerr(Ds[b].to_numpy().reshape(refs[b].shape), 1.25 * refs[b]) <= TOL, (name, b)


def test_sliced_operands_in_place(sip, oracle):
    """`T[a,i,mu,j] = T2[a,i,b,j] * ca[mu,b]` with ca read in place from the static array (strided operand) and a strided
    destination: the bandwidth-shaped kernel takes strides from the parent arrays (no split of a strided destination)"""
    sip.set_tuning("lowint_scope", 2)
    rng = np.random.default_rng(77)
    T2 = rand_block(rng, (9, 4, 11, 5))
    ca = rand_block(rng, (30, 26))
    Dpar = rand_block(rng, (9, 4, 20, 5))
    dlab, llab, rlab = [1, 2, 3, 4], [1, 2, 5, 4], [3, 5]
    ptrn, ierr = sip.get_contraction_ptrn(dlab, llab, rlab)
    assert ierr == 0
    mu0, b0, d0 = 7, 13, 6
    cas = np.asfortranarray(ca[mu0:mu0 + 8, b0:b0 + 11])
    ref, e = oracle.contract_labels(dlab, [9, 4, 8, 5], llab, T2, rlab, cas)
    assert e == 0
    dT2, dca, dD = sip.DeviceBlock.from_numpy(T2), sip.DeviceBlock.from_numpy(ca), sip.DeviceBlock.from_numpy(Dpar)
    sip.contract_sliced(ptrn, dT2, [9, 4, 11, 5], None, dca, [8, 11], [mu0, b0], [9, 4, 8, 5], out=dD, dbeg=[0, 0, d0, 0])
    want = Dpar.copy(order="F")
    want[:, :, d0:d0 + 8, :] = ref
    assert relerr(dD.to_numpy(), want) <= TOL


def test_prepared_plans_replay_the_launches(sip, oracle):
    """sipgpu_plan_*: a work-list marshalled once (descriptors resident) gives the same results as the eager call, for every
    kernel family a work-list can reach (tile kernel, split-K with its pre-scale, bandwidth-shaped kernel, dot kernel), with
    different (alpha, beta) pairs on the same plan, launch after launch, and costs no more launches than the eager path."""
    rng = np.random.default_rng(11)
    for scope, (d, l, r, ext, nb) in ((0, ("aibj", "aick", "ckbj", dict(a=9, i=4, b=9, j=4, c=9, k=4), 6)),     # tiles
                                      (0, ("ab", "aicj", "bicj", dict(a=10, b=12, i=6, c=40, j=20), 2)),      # tiles, split-K
                                      (2, ("ab", "aicj", "bicj", dict(a=10, b=12, i=6, c=40, j=20), 2)),      # lowint, k slices
                                      (1, ("ab", "ac", "cb", dict(a=20, b=20, c=50), 40)),                    # lowint, tiny
                                      (1, ("ab", "cda", "cdb", dict(a=1, b=1, c=50, d=41), 5))):              # dot kernel
        sip.set_tuning("lowint_scope", scope)
        labs = sorted(set(d + l + r))
        num = {c: i + 1 for i, c in enumerate(labs)}
        dl_, ll_, rl_ = [num[c] for c in d], [num[c] for c in l], [num[c] for c in r]
        ptrn, ierr = sip.get_contraction_ptrn(dl_, ll_, rl_)
        assert ierr == 0
        lsh, rsh, dsh = tuple(ext[c] for c in l), tuple(ext[c] for c in r), tuple(ext[c] for c in d)
        Lh = [rand_block(rng, lsh) for _ in range(2 * nb)]
        Rh = [rand_block(rng, rsh) for _ in range(2 * nb)]
        Ld, Rd = [sip.DeviceBlock.from_numpy(x) for x in Lh], [sip.DeviceBlock.from_numpy(x) for x in Rh]
        refs = []
        for b in range(nb):
            acc = 0.0
            for c in (2 * b, 2 * b + 1):
                t, e = oracle.contract_labels(dl_, list(dsh), ll_, Lh[c], rl_, Rh[c])
                assert e == 0
                acc = acc + t
            refs.append(acc)
        Ds = [sip.DeviceBlock(dsh) for _ in range(nb)]
        bc = sip.BatchedContraction(ptrn, [lsh] * nb, [rsh] * nb, [dsh] * nb, [x.ptr for x in Ld], [x.ptr for x in Rd],
                                    [x.ptr for x in Ds], chain_start=[2 * b for b in range(nb + 1)])
        l0 = sip.kernel_launches()
        bc.launch(prepared=False)
        eager = sip.kernel_launches() - l0
        for b in range(nb):
            assert relerr(Ds[b].to_numpy().reshape(refs[b].shape), refs[b]) <= TOL
            Ds[b].fill(0.0)
        for rep in range(3):                       # beta = 0 replays
            l0 = sip.kernel_launches()
            bc.launch()
            assert sip.kernel_launches() - l0 == eager
            for b in range(nb):
                assert relerr(Ds[b].to_numpy().reshape(refs[b].shape), refs[b]) <= TOL
        bc.launch(alpha=-0.5, beta=1.0)            # a second (alpha, beta) variant of the same plan
        bc.launch(alpha=-0.5, beta=1.0)
        for b in range(nb):
            assert np.max(np.abs(Ds[b].to_numpy().reshape(refs[b].shape))) <= 1e-9 * max(1.0, np.max(np.abs(refs[b])))
        bc.destroy()
    sip.set_tuning("lowint_scope", 1)


# ---- the TMA-fed slab path (slab_kernel): small results of long contractions whose chunks are contiguous runs ----
SLAB_CASES = [
    ("50x50, free index fastest", "ab", "acde", "bcde", dict(a=50, b=50, c=20, d=10, e=6)),
    ("20x20, contracted index fastest", "ab", "cade", "cbde", dict(a=20, b=20, c=50, d=6, e=5)),
    ("20x50, different orders of the outer contracted indices", "ab", "acde", "cbed", dict(a=20, b=50, c=20, d=10, e=8)),
    ("50x50, one operand in many runs", "ab", "acde", "bedc", dict(a=50, b=50, c=20, d=4, e=16)),
    ("50x20", "ab", "acde", "bcde", dict(a=50, b=20, c=20, d=10, e=6)),
    ("16x16", "ab", "acde", "bcde", dict(a=16, b=16, c=16, d=16, e=8)),
    ("32x32", "ab", "acde", "bcde", dict(a=32, b=32, c=8, d=8, e=16)),
    ("40x40", "ab", "acde", "bcde", dict(a=40, b=40, c=10, d=12, e=10)),
    ("64x64", "ab", "acde", "bcde", dict(a=64, b=64, c=16, d=8, e=8)),
    ("56x24", "ab", "acde", "bcde", dict(a=56, b=24, c=12, d=10, e=10)),
    ("12x60", "ab", "acde", "bcde", dict(a=12, b=60, c=12, d=10, e=10)),
    ("odd extents (gather kernel)", "ab", "acde", "bcde", dict(a=25, b=25, c=11, d=10, e=10)),
    ("20x50, hybrid: gathered operand contracted-fastest", "ab", "acde", "cdeb", dict(a=20, b=50, c=10, d=12, e=10)),
    ("50x25, hybrid: gathered operand in 8-byte items", "ab", "acde", "bedc", dict(a=50, b=25, c=20, d=4, e=16)),
    ("20x20, hybrid: gathered operand on the left", "ab", "aedc", "bcde", dict(a=20, b=20, c=20, d=6, e=10)),
    ("rank-5 operands", "ab", "xacde", "xbcde", dict(a=20, b=20, c=10, d=10, e=6, x=2)),
]


@pytest.mark.parametrize("slab", [1, 0], ids=["slab", "gather"])
@pytest.mark.parametrize("case", SLAB_CASES, ids=[c[0] for c in SLAB_CASES])
def test_slab_shapes_single_block_split_along_k(sip, oracle, case, slab):
    """one destination: the launch is cut along K into slices that meet through red.add (beta by a pre-scale)"""
    sip.set_tuning("lowint_scope", 1)
    sip.set_tuning("lowint_slab", slab)
    try:
        name, d, l, r, ext = case
        labs = sorted(set(d + l + r))
        num = {c: i + 1 for i, c in enumerate(labs)}
        rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
        L = rand_block(rng, tuple(ext[c] for c in l))
        R = rand_block(rng, tuple(ext[c] for c in r))
        dext = [ext[c] for c in d]
        ref, ierr = oracle.contract_labels([num[c] for c in d], dext, [num[c] for c in l], L, [num[c] for c in r], R)
        assert ierr == 0
        dL, dR = sip.DeviceBlock.from_numpy(L), sip.DeviceBlock.from_numpy(R)
        got = sip.contract_labels([num[c] for c in d], dext, [num[c] for c in l], dL, [num[c] for c in r], dR).to_numpy()
        assert relerr(got.reshape(ref.shape), ref) <= TOL, name
        D0 = rand_block(rng, tuple(dext))
        out = sip.DeviceBlock.from_numpy(D0)
        sip.contract_labels([num[c] for c in d], dext, [num[c] for c in l], dL, [num[c] for c in r], dR, out=out, alpha=-0.5, beta=2.0)
        assert relerr(out.to_numpy().reshape(ref.shape), -0.5 * ref + 2.0 * D0) <= TOL, name
    finally:
        sip.set_tuning("lowint_slab", 1)


@pytest.mark.parametrize("nblocks,chain_max", [(2, 5), (37, 3), (400, 2)])
@pytest.mark.parametrize("case", SLAB_CASES[:5] + SLAB_CASES[7:9] + SLAB_CASES[12:15], ids=lambda c: c[0])
def test_slab_shapes_work_lists_with_chains(sip, oracle, case, nblocks, chain_max):
    """many destinations (one item per CTA slot, the ring running across item boundaries), chains of unequal length, the
    deterministic reduction over the four warps of the K-split variants: two launches give identical bits"""
    sip.set_tuning("lowint_scope", 1)
    sip.set_tuning("lowint_slab", 1)
    name, d, l, r, ext = case
    labs = sorted(set(d + l + r))
    num = {c: i + 1 for i, c in enumerate(labs)}
    dl_, ll_, rl_ = [num[c] for c in d], [num[c] for c in l], [num[c] for c in r]
    ptrn, ierr = sip.get_contraction_ptrn(dl_, ll_, rl_)
    assert ierr == 0
    rng = np.random.default_rng(nblocks * 31 + chain_max)
    lsh, rsh, dsh = tuple(ext[c] for c in l), tuple(ext[c] for c in r), tuple(ext[c] for c in d)
    npool = 6
    Lh = [rand_block(rng, lsh) for _ in range(npool)]
    Rh = [rand_block(rng, rsh) for _ in range(npool)]
    Ld, Rd = [sip.DeviceBlock.from_numpy(x) for x in Lh], [sip.DeviceBlock.from_numpy(x) for x in Rh]
    prod = {}

    def pair_product(i, j):
        if (i, j) not in prod:
            prod[(i, j)], e = oracle.contract_labels(dl_, list(dsh), ll_, Lh[i], rl_, Rh[j])
            assert e == 0
        return prod[(i, j)]

    lp, rp, chain, refs, Ds = [], [], [0], [], []
    for b in range(nblocks):
        n = 1 + (b * 7) % chain_max
        ref = np.zeros(dsh, order="F")
        for c in range(n):
            i, j = (b * 3 + c) % npool, (b * 5 + 2 * c + 1) % npool
            lp.append(Ld[i].ptr), rp.append(Rd[j].ptr)
            ref = ref + pair_product(i, j)
        chain.append(len(lp))
        refs.append(ref)
        Ds.append(sip.DeviceBlock(dsh))
    bc = sip.BatchedContraction(ptrn, [lsh] * nblocks, [rsh] * nblocks, [dsh] * nblocks, lp, rp, [x.ptr for x in Ds], chain_start=chain)
    bc.launch()
    step = max(1, nblocks // 25)
    first = {}
    for b in range(0, nblocks, step):
        first[b] = Ds[b].to_numpy()
        assert relerr(first[b].reshape(refs[b].shape), refs[b]) <= TOL, (name, b)
    if nblocks >= 400:      # no split along K: bit-identical from launch to launch
        bc.launch()
        for b in range(0, nblocks, step):
            assert np.array_equal(Ds[b].to_numpy(), first[b]), (name, b)
    bc.launch(alpha=0.25, beta=1.0)
    for b in range(0, nblocks, step):
        assert relerr(Ds[b].to_numpy().reshape(refs[b].shape), 1.25 * refs[b]) <= TOL, (name, b)
