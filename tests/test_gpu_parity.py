"""GPU parity: the CUDA path (through the C ABI of include/sipgpu.h) against the CPU oracle.

All tests here need a B200 (`-m gpu`).  Integer-valued cases are compared bit-exactly (the reference's own tests
use EXPECT_DOUBLE_EQ on such data); random FP64 cases use the tolerance BASELINE.json states for blocks:
max|delta| / max|ref| <= 1e-10.
"""
import itertools
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10  # BASELINE.json north_star: "within 1e-10 relative on blocks"


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


def relerr(a, ref):
    ref = np.asarray(ref)
    a = np.asarray(a)
    assert a.shape == ref.shape
    d = np.max(np.abs(a - ref)) if ref.size else 0.0
    m = np.max(np.abs(ref)) if ref.size else 1.0
    return d / (m if m > 0 else 1.0)


def rand_block(rng, shape):
    return np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape))


# ---------------------------------------------------------------------------------------------------
# the reference's known-answer tests, through the CUDA path
# ---------------------------------------------------------------------------------------------------
def test_contraction_small_test_host_abi(sip, oracle):
    # BasicSial.contraction_small_test (test_basic_sial.cpp:695-770): c[i,l] = a[i,j,k,l]*b[j,k], segs 15
    n = 15
    a = oracle.fill_cyclic((n, n, n, n), 1.0)
    b = oracle.fill_cyclic((n, n), 1.0)
    ptrn, ierr = sip.get_contraction_ptrn([1, 4], [1, 2, 3, 4], [2, 3])
    assert ierr == 0 and ptrn == [1, -1, -2, 2, -2, -3]
    c, ierr = sip.tensor_block_contract(ptrn, a, b, [n, n])
    assert ierr == 0
    ref, oerr = oracle.contract_labels([1, 4], [n, n], [1, 2, 3, 4], a, [2, 3], b)
    assert oerr == 0
    assert np.array_equal(c, ref)
    assert np.array_equal(c, np.einsum("ijkl,jk->il", a, b))


def test_contraction_small_test2_device_abi(sip, oracle):
    # BasicSial.contraction_small_test2 (:773-815): c[mu,i1,a1,i] = b[lambda,a1]*a[mu,i1,i,lambda]
    MU, LA, I, I1, A1 = 9, 9, 5, 5, 4
    a = oracle.fill_cyclic((MU, I1, I, LA), 1.0)
    b = oracle.fill_cyclic((LA, A1), float((a.size % 20) + 1))
    mu, i1, a1, i, la = 1, 2, 3, 4, 5
    da, db = sip.DeviceBlock.from_numpy(a), sip.DeviceBlock.from_numpy(b)
    dc = sip.contract_labels([mu, i1, a1, i], [MU, I1, A1, I], [la, a1], db, [mu, i1, i, la], da)
    ref, oerr = oracle.contract_labels([mu, i1, a1, i], [MU, I1, A1, I], [la, a1], b, [mu, i1, i, la], a)
    assert oerr == 0
    assert np.array_equal(dc.to_numpy(), ref)


def test_transpose_tmp(sip, oracle):
    # BasicSial.transpose_tmp (:653-693): b[j,k,i] = a[i,j,k], sequential from 53
    a = oracle.fill_sequential((8, 8, 8), 53.0)
    db = sip.permute_labels([2, 3, 1], [1, 2, 3], sip.DeviceBlock.from_numpy(a))
    b = db.to_numpy()
    assert np.array_equal(b, oracle.permute_labels([2, 3, 1], [1, 2, 3], a))
    assert np.array_equal(b, np.transpose(a, (1, 2, 0)))


@pytest.mark.parametrize("shape", [(5, 1, 5, 1), (8, 8, 8, 8)])
def test_transpose4d_tmp(sip, oracle, shape):
    # BasicSial.transpose4d_tmp (:1285-1327) / transpose4d_square_tmp (:1329-1406): b[k,j,i,l] = a[i,j,k,l]
    a = oracle.fill_sequential(shape, 53.0) if shape[1] == 1 else oracle.fill_cyclic(shape, 1.0)
    da = sip.DeviceBlock.from_numpy(a)
    db = sip.permute_labels([3, 2, 1, 4], [1, 2, 3, 4], da)
    b = db.to_numpy()
    assert np.array_equal(b, oracle.permute_labels([3, 2, 1, 4], [1, 2, 3, 4], a))
    if shape == (8, 8, 8, 8):
        # the three scalar contractions of transpose4d_square_tmp, the last one with a permuted operand
        for (ll, lb, rl, rb) in [((1, 2, 3, 4), da, (1, 2, 3, 4), da), ((3, 2, 1, 4), db, (3, 2, 1, 4), db),
                                 ((1, 2, 3, 4), da, (3, 2, 1, 4), db)]:
            e = sip.contract_labels([], [], list(ll), lb, list(rl), rb).to_numpy()
            ref, oerr = oracle.contract_labels([], [], list(ll), lb.to_numpy(), list(rl), rb.to_numpy())
            assert oerr == 0
            assert float(e) == float(ref.ravel()[0]) == float(np.sum(a * a))


def test_contract_to_scalar(sip, oracle):
    # BasicSial.contract_to_scalar (:1037-1084): x = a[i,j]*b[i,j]
    a = oracle.fill_cyclic((8, 8), 1.0)
    b = oracle.fill_cyclic((8, 8), 5.0)
    x = sip.contract_labels([], [], [1, 2], sip.DeviceBlock.from_numpy(a), [1, 2], sip.DeviceBlock.from_numpy(b))
    assert float(x.to_numpy()) == float(np.sum(a * b))
    # same through the host ABI
    ptrn, _ = sip.get_contraction_ptrn([], [1, 2], [1, 2])
    d, ierr = sip.tensor_block_contract(ptrn, a, b, [])
    assert ierr == 0 and d[0] == np.sum(a * b)


def test_sum_op_and_scale(sip, oracle):
    # BasicSial.sum_op (:817-916): d = a + c; e = d - c;  self_multiply_test (:1111); block_scale_assign (:555)
    a = oracle.fill_sequential((20, 20), 100.0)
    c = oracle.fill_sequential((20, 20), 50.0)
    da, dc = sip.DeviceBlock.from_numpy(a), sip.DeviceBlock.from_numpy(c)
    dd = sip.DeviceBlock((20, 20)).set_add_sub(da, dc, +1.0)
    de = sip.DeviceBlock((20, 20)).set_add_sub(dd, dc, -1.0)
    assert np.array_equal(dd.to_numpy(), a + c)
    assert np.array_equal(de.to_numpy(), a)
    assert np.array_equal(da.scale(3.0).to_numpy(), a * 3.0)
    assert np.array_equal(sip.DeviceBlock((20, 20)).scale_and_copy(dc, 2.0).to_numpy(), c * 2.0)
    assert np.array_equal(sip.DeviceBlock((7, 3)).fill(42.0).to_numpy(), np.full((7, 3), 42.0))
    assert np.array_equal(dc.increment(1.5).to_numpy(), c + 1.5)
    assert np.array_equal(dc.accumulate(dd).to_numpy(), c + 1.5 + a + c)


# ---------------------------------------------------------------------------------------------------
# permutes: every rank-4 pattern, ragged and tiny extents, ranks 2..6
# ---------------------------------------------------------------------------------------------------
PERMUTE_ROUTES = {"registers": (0, -1), "tma": (1, -1), "ring": (2, 0), "ring16": (2, 1), "auto": (3, -1)}


@pytest.fixture(params=list(PERMUTE_ROUTES))
def permute_route(request, sip):
    """the three permute kernels -- the cp.async ring (LDGSTS into a 3-stage ring of staged tiles; with 8-byte and with 16-byte
    loads / stores where the runs are even and aligned), TMA bulk copies (cp.async.bulk) and the register-staged tiles -- and the
    default per-launch choice all run the same cases, so that every one stays covered"""
    sip.set_tuning("permute_bulk", PERMUTE_ROUTES[request.param][0])
    sip.set_tuning("permute_vec", PERMUTE_ROUTES[request.param][1])
    yield request.param
    sip.set_tuning("permute_bulk", PERMUTE_ROUTES["auto"][0])
    sip.set_tuning("permute_vec", PERMUTE_ROUTES["auto"][1])


@pytest.mark.parametrize("shape", [(16, 16, 16, 16), (13, 30, 50, 7), (5, 8, 9, 5), (32, 3, 1, 33), (64, 20, 2, 50), (50, 20, 50, 20),
                                   (18, 6, 34, 10), (40, 40, 40, 6)])
def test_all_rank4_permutes(sip, oracle, shape, permute_route):
    rng = np.random.default_rng(7)
    a = rand_block(rng, shape)
    da = sip.DeviceBlock.from_numpy(a)
    for perm in itertools.permutations(range(4)):
        transp = [1] + [p + 1 for p in perm]  # new position of old dim i
        out = sip.permute(da, transp).to_numpy()
        ref = oracle.block_copy(a, transp)
        assert np.array_equal(out, ref), (shape, perm)


@pytest.mark.parametrize("rank", [1, 2, 3, 5, 6])
def test_permutes_other_ranks(sip, oracle, rank, permute_route):
    rng = np.random.default_rng(rank)
    pyrng = random.Random(rank)
    for trial in range(12):
        shape = tuple(pyrng.choice([1, 2, 3, 5, 8, 11, 16, 21, 6, 12]) for _ in range(rank))
        perm = list(range(rank))
        pyrng.shuffle(perm)
        a = rand_block(rng, shape)
        transp = [1] + [p + 1 for p in perm]
        out = sip.permute(sip.DeviceBlock.from_numpy(a), transp).to_numpy()
        assert np.array_equal(out, oracle.block_copy(a, transp)), (shape, perm)


@pytest.mark.parametrize("shape", [(50, 20, 50, 20), (13, 30, 7, 9), (64, 64, 16, 3), (70, 70, 70), (66, 10, 34), (16, 16, 16, 16)])
def test_permute_batched_and_accumulate(sip, oracle, shape, permute_route):
    """n blocks per launch, plain and fused permute-accumulate (out = alpha * P(in) + beta * out), incl. tiles of
    every element-per-thread class and ragged tiles."""
    rng = np.random.default_rng(11)
    pyrng = random.Random(5)
    n = 7
    for trial in range(6):
        perm = list(range(len(shape)))
        pyrng.shuffle(perm)
        transp = [1] + [p + 1 for p in perm]
        new_shape = [0] * len(shape)
        for i, p in enumerate(perm):
            new_shape[p] = shape[i]
        ins = [rand_block(rng, shape) for _ in range(n)]
        outs0 = [rand_block(rng, new_shape) for _ in range(n)]
        d_in = [sip.DeviceBlock.from_numpy(x) for x in ins]
        d_out = [sip.DeviceBlock.from_numpy(x) for x in outs0]
        sip.permute_batched(d_in, transp, d_out)
        for x, d in zip(ins, d_out):
            assert np.array_equal(d.to_numpy(), oracle.block_copy(x, transp)), (shape, perm)
        d_out = [sip.DeviceBlock.from_numpy(x) for x in outs0]
        sip.permute_batched(d_in, transp, d_out, alpha=-0.5, beta=2.0)
        for x, o, d in zip(ins, outs0, d_out):
            ref = -0.5 * oracle.block_copy(x, transp) + 2.0 * o
            assert relerr(d.to_numpy(), ref) <= 1e-14, (shape, perm)


@pytest.mark.parametrize("shape", [(16, 16, 16, 16), (50, 20, 50, 20), (18, 6, 34, 10), (13, 30, 8, 7)])
def test_permute_of_blocks_that_are_not_16_byte_aligned(sip, oracle, shape, permute_route):
    """the 16-byte paths (vector cp.async / stores, TMA) need 16-byte aligned blocks: views that start one element into an
    allocation must take the 8-byte paths -- input only, output only, both; plain and accumulating"""
    rng = np.random.default_rng(23)
    pyrng = random.Random(9)
    n = int(np.prod(shape))
    for trial in range(4):
        perm = list(range(4))
        pyrng.shuffle(perm)
        transp = [1] + [p + 1 for p in perm]
        new_shape = [0] * 4
        for i, p in enumerate(perm):
            new_shape[p] = shape[i]
        a, o = rand_block(rng, shape), rand_block(rng, new_shape)
        for off_in, off_out in ((1, 0), (0, 1), (1, 1)):
            big_in, big_out = sip.DeviceBlock((n + 2,)), sip.DeviceBlock((n + 2,))
            vin = sip.DeviceBlock(shape, ptr=big_in.ptr + 8 * off_in, owned=False)
            vout = sip.DeviceBlock(tuple(new_shape), ptr=big_out.ptr + 8 * off_out, owned=False)
            vin.scale_and_copy(sip.DeviceBlock.from_numpy(a), 1.0)
            vout.scale_and_copy(sip.DeviceBlock.from_numpy(o), 1.0)
            sip.permute_batched([vin], transp, [vout])
            assert np.array_equal(vout.to_numpy(), oracle.block_copy(a, transp)), (shape, perm, off_in, off_out)
            sip.permute_batched([vin], transp, [vout], alpha=0.25, beta=-1.0)
            ref = 0.25 * oracle.block_copy(a, transp) - oracle.block_copy(a, transp)
            assert relerr(vout.to_numpy(), ref) <= 1e-14, (shape, perm, off_in, off_out)


def test_permute_host_abi_and_large(sip, oracle):
    rng = np.random.default_rng(3)
    a = rand_block(rng, (50, 20, 50, 20))
    for transp in ([1, 3, 4, 1, 2], [1, 2, 1, 4, 3], [1, 4, 3, 2, 1], [1, 1, 2, 3, 4]):
        out, ierr = sip.tensor_block_copy(a, transp)
        assert ierr == 0
        assert np.array_equal(out, oracle.block_copy(a, transp))


# ---------------------------------------------------------------------------------------------------
# contractions: random label patterns against the oracle
# ---------------------------------------------------------------------------------------------------
def random_pattern(pyrng, max_rank=4, exts=(1, 2, 3, 4, 5, 7, 9)):
    nfl = pyrng.randint(0, 3)
    nfr = pyrng.randint(0, 3)
    nc = pyrng.randint(0, 3)
    if nfl + nc == 0 or nfr + nc == 0:
        nc = 1
    if nfl + nc > max_rank + 2 or nfr + nc > max_rank + 2:
        nc = 1
    labels = list(range(1, nfl + nfr + nc + 1))
    ext = {lab: pyrng.choice(exts) for lab in labels}
    fl, fr, cc = labels[:nfl], labels[nfl:nfl + nfr], labels[nfl + nfr:]
    llab = fl + cc
    rlab = fr + cc
    dlab = fl + fr
    pyrng.shuffle(llab)
    pyrng.shuffle(rlab)
    pyrng.shuffle(dlab)
    return dlab, llab, rlab, ext


def check_contraction(sip, oracle, rng, dlab, llab, rlab, ext, alpha=1.0, beta=0.0):
    L = rand_block(rng, tuple(ext[x] for x in llab))
    R = rand_block(rng, tuple(ext[x] for x in rlab))
    dext = [ext[x] for x in dlab]
    ref, oerr = oracle.contract_labels(dlab, dext, llab, L, rlab, R)
    assert oerr == 0
    dL, dR = sip.DeviceBlock.from_numpy(L), sip.DeviceBlock.from_numpy(R)
    if beta != 0.0:
        D0 = rand_block(rng, tuple(dext))
        out = sip.DeviceBlock.from_numpy(D0)
        sip.contract_labels(dlab, dext, llab, dL, rlab, dR, out=out, alpha=alpha, beta=beta)
        ref = alpha * ref + beta * D0
    else:
        out = sip.contract_labels(dlab, dext, llab, dL, rlab, dR, alpha=alpha)
        ref = alpha * ref
    got = out.to_numpy().reshape(ref.shape)
    assert relerr(got, ref) <= TOL, (dlab, llab, rlab, ext)
    return got


@pytest.fixture(params=["lowint", "tiles"])
def route(request, sip):
    """small test blocks are all low-intensity: run them through the bandwidth-shaped kernel (the default route) AND forced
    through the 128-wide tile kernel, so that both stay covered"""
    sip.set_tuning("lowint_scope", 2 if request.param == "lowint" else 0)
    yield request.param
    sip.set_tuning("lowint_scope", 1)


def test_random_patterns(sip, oracle, route):
    pyrng = random.Random(1234)
    rng = np.random.default_rng(1234)
    for trial in range(150):
        dlab, llab, rlab, ext = random_pattern(pyrng)
        check_contraction(sip, oracle, rng, dlab, llab, rlab, ext)


def test_random_patterns_alpha_beta(sip, oracle, route):
    pyrng = random.Random(99)
    rng = np.random.default_rng(99)
    for trial in range(40):
        dlab, llab, rlab, ext = random_pattern(pyrng)
        check_contraction(sip, oracle, rng, dlab, llab, rlab, ext, alpha=-0.5, beta=2.0)


# The label patterns of the LCCD / CCSD doubles equations (src/sialx/qm/cc/rlccd_rhf.sialx:342-355,399-417,
# 482-556; rccsd_rhf.sialx:927-959,1556-1751) as (D, L, R) label strings: v = virtual, o = occupied extents.
SIAL_PATTERNS = [
    ("aibj", "aicj", "cb"),      # one-particle: R[a,i,b,j] = T[a,i,c,j] * F[c,b]
    ("aibj", "aibk", "kj"),
    ("aibj", "ikjl", "akbl"),    # hh ladder: T2new[a,i,b,j] = W[i,k,j,l]... (label permuted below)
    ("aibj", "akbl", "kilj"),
    ("aibj", "aick", "ckbj"),    # ph ring
    ("aibj", "akci", "ckbj"),
    ("aibj", "bkci", "akcj"),
    ("minj", "manb", "aibj"),    # AO ladder shape: Y[m,i,n,j] = V[l,m,s,n] * T[l,i,s,j]
    ("ij", "iakb", "jakb"),      # rank-2 result (7x(2,4,4) in rccsd)
    ("ab", "aibj", "ij"),        # (2,4,2)
]


@pytest.mark.parametrize("o,v", [(4, 6), (5, 9), (8, 16)])
def test_sial_cc_patterns(sip, oracle, o, v, route):
    rng = np.random.default_rng(o * 100 + v)
    for dl, ll, rl in SIAL_PATTERNS:
        labs = sorted(set(dl + ll + rl))
        num = {c: i + 1 for i, c in enumerate(labs)}
        ext = {num[c]: (v if c in "abcdef" else o) for c in labs}
        check_contraction(sip, oracle, rng, [num[c] for c in dl], [num[c] for c in ll], [num[c] for c in rl], ext)


@pytest.mark.parametrize("sizes", [dict(o=4, v=6, p=5, n=7, x=2, s=2), dict(o=20, v=24, p=9, n=13, x=1, s=3)])
def test_all_sial_patterns_golden(sip, oracle, sizes, route):
    """All 170 distinct contraction patterns of the reference's CC / (T) / EOM SIAL programs (tests/golden/
    sial_contraction_patterns.txt) through the fused kernel, against the oracle, at 1e-10."""
    from conftest import sial_patterns

    rng = np.random.default_rng(sizes["o"])
    oracle.use_openblas(8)
    try:
        for d, l, r, kinds, where in sial_patterns():
            num = {c: i + 1 for i, c in enumerate(kinds)}
            ext = {num[c]: sizes[k] for c, k in kinds.items()}
            check_contraction(sip, oracle, rng, [num[c] for c in d], [num[c] for c in l], [num[c] for c in r], ext)
    finally:
        oracle.use_naive_gemm()


@pytest.mark.parametrize("s", [16, 24])
def test_sweep_rank4_full_cross_product_sample(sip, oracle, s):
    # config 4 of BASELINE.json: D[p0..p3] = L[..]*R[..], 2 contracted indices, sampled across destination
    # permutations x placements of the contracted labels (the full cross product runs in the bench sweep)
    pyrng = random.Random(s)
    rng = np.random.default_rng(s)
    oracle.use_openblas(8)
    try:
        for trial in range(24):
            fl, fr, cc = [1, 2], [3, 4], [5, 6]
            llab, rlab, dlab = fl + cc, fr + cc, fl + fr
            pyrng.shuffle(llab), pyrng.shuffle(rlab), pyrng.shuffle(dlab)
            check_contraction(sip, oracle, rng, dlab, llab, rlab, {i: s for i in range(1, 7)})
    finally:
        oracle.use_naive_gemm()


def test_ragged_and_eom_shapes(sip, oracle, route):
    rng = np.random.default_rng(5)
    # ragged extents 13,30,50,64 (config 4) and EOM-style rank-5 blocks with a leading extent-1 index (config 3)
    check_contraction(sip, oracle, rng, [1, 2, 3, 4], [1, 5, 2, 6], [6, 3, 5, 4], {1: 13, 2: 30, 3: 50, 4: 7, 5: 9, 6: 11})
    check_contraction(sip, oracle, rng, [7, 1, 2, 3, 4], [7, 1, 5, 2, 6], [6, 3, 5, 4],
                      {1: 8, 2: 5, 3: 8, 4: 5, 5: 8, 6: 5, 7: 1})
    check_contraction(sip, oracle, rng, [1, 2], [3, 1, 4, 2, 5], [3, 4, 5], {1: 8, 2: 5, 3: 1, 4: 8, 5: 5})
    # outer product (no contracted index) and matrix-vector shapes
    check_contraction(sip, oracle, rng, [1, 2, 3], [1, 3], [2], {1: 17, 2: 9, 3: 4})
    check_contraction(sip, oracle, rng, [1], [1, 2, 3], [3, 2], {1: 33, 2: 12, 3: 7})


def test_contract_sliced_static_arrays(sip, oracle):
    """Operands that are blocks of a static (contiguous) array, read in place: the reference extracts the block
    (tensor_block_slice_, F90:271-330), contracts and frees it; the fused kernel reads the parent array through its
    strides.  Half transformation T[a,i,mu,j] = T2[a,i,b,j] * ca[mu,b] with ca a (norb x nmo) static array, and a
    destination written in place as a slice (insert fused as well)."""
    rng = np.random.default_rng(21)
    norb, nmo, v, o = 37, 44, 12, 5
    ca = rand_block(rng, (norb, nmo))
    T2 = rand_block(rng, (v, o, v, o))
    dca, dT2 = sip.DeviceBlock.from_numpy(ca), sip.DeviceBlock.from_numpy(T2)
    ptrn, ierr = sip.get_contraction_ptrn([1, 2, 5, 4], [1, 2, 3, 4], [5, 3])
    assert ierr == 0
    for (mu0, nmu, b0) in ((0, 13, 5), (13, 13, 17), (26, 11, 32), (1, 12, 20)):  # even / odd offsets and extents
        ca_blk, ierr = oracle.block_slice(ca, (nmu, v), (mu0, b0))  # what the reference does first (block.cpp:272-297)
        assert ierr == 0 and np.array_equal(ca_blk, ca[mu0:mu0 + nmu, b0:b0 + v])
        ref, ierr = oracle.block_contract(ptrn, T2, ca_blk, (v, o, nmu, o))
        assert ierr == 0
        got = sip.contract_sliced(ptrn, dT2, (v, o, v, o), None, dca, (nmu, v), (mu0, b0), (v, o, nmu, o))
        assert relerr(got.to_numpy(), ref) <= 1e-10
        # destination as a slice of a larger array [v, o, norb, o], accumulate form
        big = rand_block(rng, (v, o, norb, o))
        dbig = sip.DeviceBlock.from_numpy(big)
        sip.contract_sliced(ptrn, dT2, (v, o, v, o), None, dca, (nmu, v), (mu0, b0), (v, o, nmu, o), out=dbig,
                            dbeg=(0, 0, mu0, 0), alpha=0.5, beta=1.0)
        want = big.copy(order="F")
        want[:, :, mu0:mu0 + nmu, :] += 0.5 * ref
        assert relerr(dbig.to_numpy(), want) <= 1e-10
    # a slice that does not fit its parent is an argument error, not a wild read
    with pytest.raises(sip.SipGpuError):
        sip.contract_sliced(ptrn, dT2, (v, o, v, o), None, dca, (13, v), (30, 0), (v, o, 13, o))


def test_scalar_operand_cases(sip, oracle):
    # F90:764-780: rank-0 operands
    rng = np.random.default_rng(11)
    T = rand_block(rng, (6, 5, 4))
    sc = np.array(0.75)
    dT, dS = sip.DeviceBlock.from_numpy(T), sip.DeviceBlock.from_numpy(sc)
    out = sip.contract_labels([2, 3, 1], [5, 4, 6], [1, 2, 3], dT, [], dS).to_numpy()
    assert relerr(out, np.transpose(T, (1, 2, 0)) * 0.75) <= TOL
    out = sip.contract_labels([1, 2, 3], [6, 5, 4], [], dS, [1, 2, 3], dT).to_numpy()
    assert relerr(out, T * 0.75) <= TOL
    s2 = sip.contract_labels([], [], [], dS, [], dS).to_numpy()
    assert float(s2) == 0.75 * 0.75


def test_illegal_patterns(sip):
    # get_contraction_ptrn error codes (F90:106-131)
    assert sip.get_contraction_ptrn([1, 2], [1, 3], [3])[1] == 2
    assert sip.get_contraction_ptrn([1, 2], [1, 3], [3, 4])[1] == 4
    assert sip.get_contraction_ptrn([1], [1, 1], [1])[1] == 5
    assert sip.get_contraction_ptrn([1, 1], [2], [2])[1] == 6
    # extents that disagree with the pattern -> ierr 1 from contract (contr_ptrn_ok)
    a = np.zeros((3, 4), order="F")
    b = np.zeros((5, 6), order="F")
    _, ierr = sip.tensor_block_contract([1, -1, -2, 2], a, b, [3, 6])
    assert ierr == 1
    with pytest.raises(sip.SipGpuError):
        sip.contract_labels([1, 2], [3, 6], [1, 3], sip.DeviceBlock((3, 4)), [4, 2], sip.DeviceBlock((5, 6)))


# ---------------------------------------------------------------------------------------------------
# batched work-lists, raw GEMM view, elementwise host ABI
# ---------------------------------------------------------------------------------------------------
def test_batched_heterogeneous(sip, oracle):
    rng = np.random.default_rng(21)
    pyrng = random.Random(21)
    # R[a,i,b,j] = T[a,i,c,k] * V[c,k,b,j] over blocks with ragged segment extents
    dlab, llab, rlab = [1, 2, 3, 4], [1, 2, 5, 6], [5, 6, 3, 4]
    ptrn, ierr = sip.get_contraction_ptrn(dlab, llab, rlab)
    assert ierr == 0
    Ls, Rs, Ds, refs = [], [], [], []
    for p in range(37):
        e = {k: pyrng.choice([3, 5, 8, 9, 16]) for k in range(1, 7)}
        L = rand_block(rng, tuple(e[x] for x in llab))
        R = rand_block(rng, tuple(e[x] for x in rlab))
        ref, oerr = oracle.contract_labels(dlab, [e[x] for x in dlab], llab, L, rlab, R)
        assert oerr == 0
        Ls.append(sip.DeviceBlock.from_numpy(L))
        Rs.append(sip.DeviceBlock.from_numpy(R))
        Ds.append(sip.DeviceBlock([e[x] for x in dlab]))
        refs.append(ref)
    sip.set_tuning("lowint_scope", 0)   # the tile kernel: heterogeneous extents share a launch
    before = sip.kernel_launches()
    sip.contract_batched(ptrn, Ls, Rs, Ds)
    assert sip.kernel_launches() - before <= 4  # one launch per kernel variant, not per block
    sip.set_tuning("lowint_scope", 2)    # the bandwidth-shaped kernel: one launch per distinct shape
    Ds2 = [sip.DeviceBlock(d.shape) for d in Ds]
    before = sip.kernel_launches()
    sip.contract_batched(ptrn, Ls, Rs, Ds2)
    assert sip.kernel_launches() - before <= len({(l.shape, r.shape) for l, r in zip(Ls, Rs)})
    for d, ref in zip(Ds2, refs):
        assert relerr(d.to_numpy(), ref) <= TOL
    sip.set_tuning("lowint_scope", 1)
    for d, ref in zip(Ds, refs):
        assert relerr(d.to_numpy(), ref) <= TOL
    # fused accumulate over the same work-list: D = D + L*R
    sip.contract_batched(ptrn, Ls, Rs, Ds, alpha=1.0, beta=1.0)
    for d, ref in zip(Ds, refs):
        assert relerr(d.to_numpy(), 2.0 * ref) <= TOL


def test_chained_block_sparse(sip, oracle, route):
    # hh-ladder body (rlccd_rhf.sialx:342-355): T2new[a,i,b,j] += T2old[a,i1,b,j1] * V[i,i1,j,j1] summed over the
    # (i1,j1) SEGMENTS inside one launch; ragged segment extents per destination
    rng = np.random.default_rng(31)
    pyrng = random.Random(31)
    dlab, llab, rlab = [1, 2, 3, 4], [1, 5, 3, 6], [2, 5, 4, 6]
    ptrn, ierr = sip.get_contraction_ptrn(dlab, llab, rlab)
    assert ierr == 0
    lsh, rsh, dsh, lp, rp, dp, chain, refs, keep = [], [], [], [], [], [], [0], [], []
    for dest in range(11):
        e = {k: pyrng.choice([4, 6, 9, 16]) for k in range(1, 7)}
        nchain = pyrng.randint(1, 5)
        ref = np.zeros([e[x] for x in dlab], order="F")
        for c in range(nchain):
            L = rand_block(rng, tuple(e[x] for x in llab))
            R = rand_block(rng, tuple(e[x] for x in rlab))
            r1, oerr = oracle.contract_labels(dlab, [e[x] for x in dlab], llab, L, rlab, R)
            assert oerr == 0
            ref += r1
            dl, dr = sip.DeviceBlock.from_numpy(L), sip.DeviceBlock.from_numpy(R)
            keep += [dl, dr]
            lp.append(dl.ptr), rp.append(dr.ptr)
        chain.append(len(lp))
        d = sip.DeviceBlock([e[x] for x in dlab])
        keep.append(d)
        dp.append(d)
        lsh.append([e[x] for x in llab]), rsh.append([e[x] for x in rlab]), dsh.append([e[x] for x in dlab])
        refs.append(ref)
    bc = sip.BatchedContraction(ptrn, lsh, rsh, dsh, lp, rp, [d.ptr for d in dp], chain_start=chain)
    bc.launch()
    for d, ref in zip(dp, refs):
        assert relerr(d.to_numpy(), ref) <= TOL
    bc.launch(alpha=0.5, beta=1.0)
    for d, ref in zip(dp, refs):
        assert relerr(d.to_numpy(), 1.5 * ref) <= TOL


def test_long_contracted_index_windows(sip):
    # K larger than the kernel's k-offset window (2048): several windows per tile, ragged tail
    rng = np.random.default_rng(41)
    L = rand_block(rng, (70, 75, 9))    # [c1, c2, m]  K = 5250
    R = rand_block(rng, (75, 11, 70))   # [c2, n, c1]
    out = sip.contract_labels([3, 4], [9, 11], [1, 2, 3], sip.DeviceBlock.from_numpy(L), [2, 4, 1],
                              sip.DeviceBlock.from_numpy(R)).to_numpy()
    assert relerr(out, np.einsum("abm,bna->mn", L, R)) <= TOL


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (200, 77, 333), (1, 500, 64), (1000, 1000, 1000), (129, 257, 17), (64, 64, 5000)])
def test_dgemm_tn_view(sip, m, n, k):
    rng = np.random.default_rng(m + n + k)
    A = rand_block(rng, (k, m))
    B = rand_block(rng, (k, n))
    dA, dB, dC = sip.DeviceBlock.from_numpy(A), sip.DeviceBlock.from_numpy(B), sip.DeviceBlock((m, n))
    sip.dgemm_tn(m, n, k, dA, k, dB, k, dC, m)
    assert relerr(dC.to_numpy(), A.T @ B) <= TOL


def test_host_abi_elementwise(sip, oracle):
    rng = np.random.default_rng(8)
    t = rand_block(rng, (9, 7, 5))
    u = rand_block(rng, (9, 7, 5))
    out, ierr = sip.tensor_block_add(t, u, 0.25)
    ref, _ = oracle.block_add(t.copy(order="F"), u, 0.25)
    assert ierr == 0 and relerr(out, ref) <= 1e-15
    out, ierr = sip.tensor_block_scale(t, -3.0)
    assert ierr == 0 and np.array_equal(out, t * -3.0)
    out, ierr = sip.tensor_block_init((4, 3, 2), 7.5)
    assert ierr == 0 and np.array_equal(out, np.full((4, 3, 2), 7.5))
    v, ierr = sip.tensor_block_norm2(t)
    assert ierr == 0 and abs(v - oracle.block_norm2(t)) <= 1e-12 * abs(v)
    s, ierr = sip.tensor_block_slice(t, [4, 3, 2], [2, 1, 3])
    sref, _ = oracle.block_slice(t, [4, 3, 2], [2, 1, 3])
    assert ierr == 0 and np.array_equal(s, sref)
    ins, ierr = sip.tensor_block_insert(t, s * 2.0, [5, 4, 0])
    iref, _ = oracle.block_insert(t.copy(order="F"), s * 2.0, [5, 4, 0])
    assert ierr == 0 and np.array_equal(ins, iref)
    n, ierr = sip.tensor_size_by_shape([3, 4, 5])
    assert (n, ierr) == (60, 0)


def test_gpu_legacy_abi(sip, oracle):
    # boundary 2: the _gpu_* entry points of gpu_super_instructions.h with label arguments
    import ctypes as C

    L = sip.lib()
    rng = np.random.default_rng(2)
    a = rand_block(rng, (6, 5, 4, 3))
    b = rand_block(rng, (4, 5, 7))
    n_a, n_b = a.size, b.size
    ga, gb = L._gpu_allocate(n_a), L._gpu_allocate(n_b)
    assert ga and gb
    assert L._gpu_host_to_device(a.ctypes.data_as(C.c_void_p), ga, n_a) == 0
    assert L._gpu_host_to_device(b.ctypes.data_as(C.c_void_p), gb, n_b) == 0
    # y[i,l,m] = a[i,j,k,l] * b[k,j,m]
    gy = L._gpu_allocate(6 * 3 * 7)
    ia = sip._ia
    assert L._gpu_contract(gy, 3, ia([6, 3, 7]), ia([1, 4, 5]), ga, 4, ia([6, 5, 4, 3]), ia([1, 2, 3, 4]), gb, 3,
                           ia([4, 5, 7]), ia([3, 2, 5])) == 0
    y = np.empty((6, 3, 7), order="F")
    assert L._gpu_device_to_host(y.ctypes.data_as(C.c_void_p), gy, y.size) == 0
    ref, _ = oracle.contract_labels([1, 4, 5], [6, 3, 7], [1, 2, 3, 4], a, [3, 2, 5], b)
    assert relerr(y, ref) <= TOL
    # permute, axpy, selfmultiply, memset, d2d
    gp = L._gpu_allocate(n_a)
    assert L._gpu_permute(gp, 4, ia([3, 4, 5, 6]), ia([4, 3, 2, 1]), ga, 4, ia([6, 5, 4, 3]), ia([1, 2, 3, 4])) == 0
    p = np.empty((3, 4, 5, 6), order="F")
    L._gpu_device_to_host(p.ctypes.data_as(C.c_void_p), gp, p.size)
    assert np.array_equal(p, np.transpose(a, (3, 2, 1, 0)))
    gz = L._gpu_allocate(n_a)  # zero-filled
    assert L._gpu_axpy(gz, ga, 2.0, n_a) == 0 and L._gpu_selfmultiply(gz, 0.5, n_a) == 0
    z = np.empty_like(a)
    L._gpu_device_to_host(z.ctypes.data_as(C.c_void_p), gz, n_a)
    assert np.array_equal(z, a)
    assert L._gpu_double_memset(gz, 3.25, n_a) == 0 and L._gpu_device_to_device(gp, gz, n_a) == 0
    L._gpu_device_to_host(z.ctypes.data_as(C.c_void_p), gp, n_a)
    assert np.all(z == 3.25)
    for g in (ga, gb, gy, gp, gz):
        assert L._gpu_free(g) == 0


# ---------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full block sizes (no oracle needed)
# ---------------------------------------------------------------------------------------------------
def test_full_size_properties(sip):
    rng = np.random.default_rng(17)
    v, o = 50, 20
    T = rand_block(rng, (v, o, v, o))
    V = rand_block(rng, (v, o, v, o))
    dT, dV = sip.DeviceBlock.from_numpy(T), sip.DeviceBlock.from_numpy(V)
    # permute round trip is the identity; permute preserves the norm
    p = sip.permute_labels([3, 4, 1, 2], [1, 2, 3, 4], dT)
    back = sip.permute_labels([1, 2, 3, 4], [3, 4, 1, 2], p)
    assert np.array_equal(back.to_numpy(), T)
    assert abs(p.norm2() - dT.norm2()) <= 1e-12 * dT.norm2()
    # ring term Z[a,i,b,j] = T[a,i,c,k] V[c,k,b,j]: compare with a BLAS matmul on the host
    Z = sip.contract_labels([1, 2, 3, 4], [v, o, v, o], [1, 2, 5, 6], dT, [5, 6, 3, 4], dV)
    ref = (T.reshape(v * o, v * o, order="F") @ V.reshape(v * o, v * o, order="F")).reshape((v, o, v, o), order="F")
    assert relerr(Z.to_numpy(), ref) <= TOL
    # linearity: contract(T, 2V) == 2 contract(T, V) exactly (scaling by 2 is exact in FP64)
    dV2 = sip.DeviceBlock((v, o, v, o)).scale_and_copy(dV, 2.0)
    Z2 = sip.contract_labels([1, 2, 3, 4], [v, o, v, o], [1, 2, 5, 6], dT, [5, 6, 3, 4], dV2)
    assert np.array_equal(Z2.to_numpy(), 2.0 * Z.to_numpy())
    # permutation equivariance: contracting permuted operands gives the permuted result, bit for bit up to
    # summation order -> within TOL
    Tp = sip.permute_labels([6, 1, 5, 2], [1, 2, 5, 6], dT)  # T'[k,a,c,i]
    Zp = sip.contract_labels([3, 1, 4, 2], [v, v, o, o], [6, 1, 5, 2], Tp, [5, 6, 3, 4], dV)
    assert relerr(Zp.to_numpy(), np.transpose(ref, (2, 0, 3, 1))) <= TOL


@pytest.mark.parametrize("case", ["single_windows", "single_ragged", "chain_pairs", "chain_windows"])
def test_split_k_partial_sums(sip, oracle, case, route):
    """few small destinations with a long contracted range (D[a,b] = L[a,i,c,j]*R[b,i,c,j], the (2,4,4) SIAL patterns):
    the launcher cuts the chain / the k windows into partial problems that meet in D through red.add (abi.cu
    run_worklist); alpha and beta must come out exactly as in the unsplit op."""
    rng = np.random.default_rng(17)
    dl, ll, rl = [1, 2], [1, 3, 4, 5], [2, 3, 4, 5]
    if case == "single_windows":
        shp, ndest, npair = (12, 20, 50, 20), 1, 1          # K = 20000: 10 k windows
    elif case == "single_ragged":
        shp, ndest, npair = (9, 13, 50, 13), 1, 1           # K = 8450: 4 full windows + 258
    elif case == "chain_pairs":
        shp, ndest, npair = (10, 10, 50, 10), 3, 24         # K = 5000 per pair, 24 pairs per destination
    else:
        shp, ndest, npair = (10, 20, 50, 20), 2, 2          # 2 pairs x 10 windows
    a = shp[0]
    alpha, beta = -0.75, 0.5
    Ls = [rng.uniform(-1, 1, shp) for _ in range(ndest * npair)]
    Rs = [rng.uniform(-1, 1, shp) for _ in range(ndest * npair)]
    D0 = [rng.uniform(-1, 1, (a, a)) for _ in range(ndest)]
    want = []
    for d in range(ndest):
        acc = np.zeros((a, a), order="F")
        for p in range(npair):
            t, ierr = oracle.contract_labels(dl, [a, a], ll, Ls[d * npair + p], rl, Rs[d * npair + p])
            assert ierr == 0
            acc += t
        want.append(alpha * acc + beta * D0[d])
    dL = [sip.DeviceBlock.from_numpy(x) for x in Ls]
    dR = [sip.DeviceBlock.from_numpy(x) for x in Rs]
    dD = [sip.DeviceBlock.from_numpy(x) for x in D0]
    ptrn, _ = sip.get_contraction_ptrn(dl, ll, rl)
    if ndest == 1 and npair == 1:
        sip.contract(ptrn, dL[0], dR[0], (a, a), out=dD[0], alpha=alpha, beta=beta)
    else:
        bc = sip.BatchedContraction(ptrn, [shp] * ndest, [shp] * ndest, [(a, a)] * ndest, [b.ptr for b in dL], [b.ptr for b in dR],
                                    [b.ptr for b in dD], chain_start=[d * npair for d in range(ndest + 1)])
        bc.launch(alpha=alpha, beta=beta)
    for d in range(ndest):
        got = dD[d].to_numpy()
        assert np.max(np.abs(got - want[d])) <= 1e-10 * np.max(np.abs(want[d]))
