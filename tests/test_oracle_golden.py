"""The CPU oracle against the reference's own known-answer tests.

Each test restates one reference test: same SIAL statement, same fills, same expected formula,
exact comparison (all values are small integers, so FP64 sums are exact -- the reference uses
EXPECT_DOUBLE_EQ / Fortran .ne.).
"""
import itertools

import numpy as np
import pytest


def test_get_contraction_ptrn_worked_example(oracle):
    # SURVEY 3.2 worked example, c[i,l] = a[i,j,k,l]*b[j,k]  (tensor_dil_omp.F90:87-142)
    i, j, k, l = 1, 2, 3, 4
    ptrn, ierr = oracle.get_contraction_ptrn([i, l], [i, j, k, l], [j, k])
    assert ierr == 0
    assert ptrn == [1, -1, -2, 2, -2, -3]
    lo2n, ro2n, do2n, dims, tr = oracle.determine_index_permutations(ptrn, [3, 4, 5, 6], [4, 5], [3, 6])
    assert lo2n == [3, 1, 2, 4] and ro2n == [1, 2] and do2n == [1, 2]
    assert dims == [18, 1, 20] and tr == [True, False, False]


def test_get_contraction_ptrn_errors(oracle):
    # F90:106-131: every label exactly twice, never twice in the same operand pair D/D
    assert oracle.get_contraction_ptrn([1, 2], [1, 3], [3])[1] == 2  # odd total
    assert oracle.get_contraction_ptrn([1, 2], [1, 3], [3, 4])[1] == 4  # unpaired label
    assert oracle.get_contraction_ptrn([1], [1, 1], [1])[1] == 5  # label more than twice
    assert oracle.get_contraction_ptrn([1, 1], [2], [2])[1] == 6  # label twice in D


def test_contraction_small_test(oracle):
    # BasicSial.contraction_small_test (test_basic_sial.cpp:695-770): c[i,l] = a[i,j,k,l]*b[j,k], segs 15
    n = 15
    a = oracle.fill_cyclic((n, n, n, n), 1.0)
    b = oracle.fill_cyclic((n, n), 1.0)
    c, ierr = oracle.contract_labels([1, 4], [n, n], [1, 2, 3, 4], a, [2, 3], b)
    assert ierr == 0
    # literal restatement of the C++ reference loop (C arrays, row-major)
    ca = ((np.arange(n ** 4) % 20) + 1).astype(np.float64).reshape(n, n, n, n)
    cb = ((np.arange(n ** 2) % 20) + 1).astype(np.float64).reshape(n, n)
    cc = np.zeros((n, n))
    for i in range(n):
        for l in range(n):
            cc[l][i] = np.sum(ca[i, :, :, l] * cb)
    c_data = c.ravel(order="F")
    for i in range(n):
        for l in range(n):
            assert cc[l][i] == c_data[i * n + l]
    # and the plain mathematical statement on column-major blocks
    assert np.array_equal(c, np.einsum("ijkl,jk->il", a, b))


def test_contraction_small_test2(oracle):
    # BasicSial.contraction_small_test2 (test_basic_sial.cpp:773-815) + test/test_contraction_small2.F:
    # c[mu,i1,a1,i] = b[lambda,a1]*a[mu,i1,i,lambda]; ao 9, occ 5, virt 4; fill counter continues a -> b
    MU, LA, I, I1, A1 = 9, 9, 5, 5, 4
    a = oracle.fill_cyclic((MU, I1, I, LA), 1.0)
    nxt = (a.size % 20) + 1
    b = oracle.fill_cyclic((LA, A1), float(nxt))
    mu, i1, a1, i, la = 1, 2, 3, 4, 5
    c, ierr = oracle.contract_labels([mu, i1, a1, i], [MU, I1, A1, I], [la, a1], b, [mu, i1, i, la], a)
    assert ierr == 0
    ref = np.zeros((MU, I1, A1, I))
    for m_ in range(MU):
        for x in range(I1):
            for y in range(I):
                for z in range(A1):
                    ref[m_, x, z, y] = sum(a[m_, x, y, q] * b[q, z] for q in range(LA))
    assert np.array_equal(c, ref)


def test_transpose_tmp(oracle):
    # BasicSial.transpose_tmp (test_basic_sial.cpp:653-693) + test/test_transpose_op.F: b[j,k,i] = a[i,j,k]
    a = oracle.fill_sequential((8, 8, 8), 53.0)
    b = oracle.permute_labels([2, 3, 1], [1, 2, 3], a)
    for i, j, k in itertools.product(range(8), repeat=3):
        assert b[j, k, i] == a[i, j, k]
    assert a[0, 0, 0] == 53.0 and a[1, 0, 0] == 54.0 and a[7, 7, 7] == 53.0 + 511


def test_transpose4d_tmp(oracle):
    # BasicSial.transpose4d_tmp (test_basic_sial.cpp:1285-1327) + test_transpose4d_op.F: b[k,j,i,l]=a[i,j,k,l]
    for shape in [(5, 1, 5, 1), (5, 5, 5, 5), (1, 5, 1, 5)]:
        a = oracle.fill_sequential(shape, 53.0)
        b = oracle.permute_labels([3, 2, 1, 4], [1, 2, 3, 4], a)
        assert np.array_equal(b, np.transpose(a, (2, 1, 0, 3)))


def test_transpose4d_square_tmp(oracle):
    # BasicSial.transpose4d_square_tmp (:1329-1406): 8^4, esum1=a*a, esum2=b*b, esum3=a[i,j,k,l]*b[k,j,i,l]
    a = oracle.fill_cyclic((8, 8, 8, 8), 1.0)
    b = oracle.permute_labels([3, 2, 1, 4], [1, 2, 3, 4], a)
    e1, ierr1 = oracle.contract_labels([], [], [1, 2, 3, 4], a, [1, 2, 3, 4], a)
    e2, ierr2 = oracle.contract_labels([], [], [3, 2, 1, 4], b, [3, 2, 1, 4], b)
    e3, ierr3 = oracle.contract_labels([], [], [1, 2, 3, 4], a, [3, 2, 1, 4], b)
    assert ierr1 == ierr2 == ierr3 == 0
    # closed form: 4096 = 204*20 + 16 ; sum_{1..20} v^2 = 2870 ; sum_{1..16} v^2 = 1496
    assert e1[0] == e2[0] == e3[0] == 204 * 2870 + 1496 == 586976.0


def test_contract_to_scalar(oracle):
    # BasicSial.contract_to_scalar (:1037-1084): x = a[i,j]*b[i,j]; a cyclic from 1, b cyclic from 5
    a = oracle.fill_cyclic((8, 8), 1.0)
    b = oracle.fill_cyclic((8, 8), 5.0)
    x, ierr = oracle.contract_labels([], [], [1, 2], a, [1, 2], b)
    ref = sum((((c % 20) + 1) * (((c + 4) % 20) + 1)) for c in range(64))
    assert ierr == 0 and x[0] == float(ref)


def test_sum_op(oracle):
    # BasicSial.sum_op (:817-916): d = a + c ; e = d - c ; 20x20, sequential from 100 / 50
    import ctypes as C

    a = oracle.fill_sequential((20, 20), 100.0)
    c = oracle.fill_sequential((20, 20), 50.0)
    d = np.empty_like(a)
    e = np.empty_like(a)
    L = oracle.lib()
    L.oracle_block_add_sub(oracle._dp(d), oracle._dp(a), oracle._dp(c), C.c_longlong(400), C.c_double(1.0))
    L.oracle_block_add_sub(oracle._dp(e), oracle._dp(d), oracle._dp(c), C.c_longlong(400), C.c_double(-1.0))
    n = np.arange(400).reshape((20, 20), order="F")
    assert np.array_equal(d, 150.0 + 2 * n) and np.array_equal(e, a)


def test_scale_fill_accumulate(oracle):
    # self_multiply_test (:1111), block_scale_assign (:555), put_accumulate in the sequential build
    import ctypes as C

    L = oracle.lib()
    a = oracle.fill_sequential((4, 5, 3), 1.0)
    b = a.copy(order="F")
    L.oracle_block_scale(oracle._dp(b), C.c_longlong(b.size), C.c_double(3.0))
    assert np.array_equal(b, 3.0 * a)
    L.oracle_block_scale_and_copy(oracle._dp(b), oracle._dp(a), C.c_longlong(b.size), C.c_double(-2.0))
    assert np.array_equal(b, -2.0 * a)
    L.oracle_block_accumulate(oracle._dp(b), oracle._dp(a), C.c_longlong(b.size))
    assert np.array_equal(b, -a)
    L.oracle_block_fill(oracle._dp(b), C.c_longlong(b.size), C.c_double(42.0))
    L.oracle_block_increment(oracle._dp(b), C.c_longlong(b.size), C.c_double(0.5))
    assert np.all(b == 42.5)
    assert oracle.block_norm2(a) == float(np.sum(a * a))


def test_put_test_closed_form(oracle):
    # Sial.put_test (test_sial.cpp:282-318): block (i,j) of segs {2,3,2} filled with k=(i-1)*3+j, then the
    # self-contraction of each block is k^2 * seg_i * seg_j
    segs = [2, 3, 2]
    for i in range(1, 4):
        for j in range(1, 4):
            k = (i - 1) * 3 + j
            blk = np.full((segs[i - 1], segs[j - 1]), float(k), order="F")
            x, ierr = oracle.contract_labels([], [], [1, 2], blk, [1, 2], blk)
            assert ierr == 0 and x[0] == k * k * segs[i - 1] * segs[j - 1]


def test_put_accumulate_closed_forms(oracle):
    # Sial.put_accumulate_mpi (:583): b=a; c=0; c+=a twice; b+c = 126 with a=42
    # Sial.put_accumulate_stress (:1072-1113): 20 iterations of c[i,j] += a,aa,a,aa with a=i, aa=j -> 20*(2i+2j)
    import ctypes as C

    L = oracle.lib()
    a = np.full((2, 3), 42.0, order="F")
    b = a.copy(order="F")
    c = np.zeros((2, 3), order="F")
    for _ in range(2):
        L.oracle_block_accumulate(oracle._dp(c), oracle._dp(a), C.c_longlong(6))
    assert np.all(b + c == 126.0)
    segs = [2, 3, 2, 2]
    for i in range(1, 5):
        for j in range(1, 5):
            cij = np.zeros((segs[i - 1], segs[j - 1]), order="F")
            ai = np.full_like(cij, float(i))
            aj = np.full_like(cij, float(j))
            for _ in range(20):
                for src in (ai, aj, ai, aj):
                    L.oracle_block_accumulate(oracle._dp(cij), oracle._dp(src), C.c_longlong(cij.size))
            assert np.all(cij == 20.0 * (2 * i + 2 * j))


def test_check_block_number_calc(oracle):
    # Sip.check_block_number_calc (test_sial.cpp:1160) / data_distribution.cpp:19-37: id -> number -> id
    nseg, lower = [3, 12, 3, 12], [1, 4, 1, 4]
    seen = set()
    for idx in itertools.product(*[range(lo, lo + n) for lo, n in zip(lower, nseg)]):
        num = oracle.block_number(nseg, lower, idx)
        assert oracle.block_num2id(nseg, lower, num) == list(idx)
        seen.add(num)
    assert seen == set(range(3 * 12 * 3 * 12))
    # last index fastest (array_table.cpp:56-66)
    assert oracle.block_number(nseg, lower, [1, 4, 1, 5]) == 1
    assert oracle.block_number(nseg, lower, [2, 4, 1, 4]) == 12 * 3 * 12
    assert oracle.block_owner(13, 8) == 5


def test_slice_insert(oracle):
    # tensor_block_slice_/insert_ (F90:271-392) as used by contiguous arrays (block.cpp:272-323)
    t = oracle.fill_sequential((7, 6, 5), 1.0)
    s, ierr = oracle.block_slice(t, (3, 2, 4), (2, 3, 1))
    assert ierr == 0 and np.array_equal(s, t[2:5, 3:5, 1:5])
    t2, ierr = oracle.block_insert(np.zeros((7, 6, 5), order="F"), s, (2, 3, 1))
    ref = np.zeros((7, 6, 5))
    ref[2:5, 3:5, 1:5] = s
    assert ierr == 0 and np.array_equal(t2, ref)


def test_add_scaled(oracle):
    # tensor_block_add_ (F90:394-436)
    t0 = oracle.fill_sequential((4, 4), 1.0)
    t1 = oracle.fill_sequential((4, 4), 10.0)
    out, ierr = oracle.block_add(t0.copy(order="F"), t1, 2.0)
    assert ierr == 0 and np.array_equal(out, t0 + 2.0 * t1)
    out, ierr = oracle.block_add(t0.copy(order="F"), t1, 1.0)
    assert np.array_equal(out, t0 + t1)


def _rand_pattern(rng, drank, ncon, max_ext=6):
    """random legal contraction: returns labels + extents"""
    nl_free = rng.integers(0, drank + 1)
    labels = list(range(1, drank + ncon + 1))
    ext = {lab: int(rng.integers(1, max_ext + 1)) for lab in labels}
    dlab = labels[:drank]
    con = labels[drank:]
    perm = list(rng.permutation(dlab))
    llab = list(rng.permutation(perm[:nl_free] + con))
    rlab = list(rng.permutation(perm[nl_free:] + con))
    return [int(x) for x in dlab], [int(x) for x in llab], [int(x) for x in rlab], ext


@pytest.mark.parametrize("seed", range(40))
def test_random_patterns_vs_einsum(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    drank = int(rng.integers(0, 5))
    ncon = int(rng.integers(0 if drank else 1, 4))
    dlab, llab, rlab, ext = _rand_pattern(rng, drank, ncon)
    if not llab or not rlab:
        pytest.skip("scalar operand case covered separately")
    L = np.asfortranarray(rng.uniform(-1, 1, [ext[x] for x in llab]))
    R = np.asfortranarray(rng.uniform(-1, 1, [ext[x] for x in rlab]))
    D, ierr = oracle.contract_labels(dlab, [ext[x] for x in dlab], llab, L, rlab, R)
    assert ierr == 0
    letters = {lab: chr(ord("a") + lab) for lab in ext}
    spec = "".join(letters[x] for x in llab) + "," + "".join(letters[x] for x in rlab) + "->" + "".join(
        letters[x] for x in dlab)
    ref = np.einsum(spec, L, R)
    got = D if drank else D[0]
    assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_scalar_operand_cases(oracle):
    # F90:771-777: tensor = tensor*scalar, scalar = scalar*scalar, with a destination permutation
    L = oracle.fill_sequential((3, 4), 1.0)
    s = np.array([2.5])
    ptrn = [2, 1]  # D[j,i] = L[i,j] * s
    D, ierr = oracle.block_contract(ptrn, L, s.reshape(()), [4, 3])
    assert ierr == 0 and np.array_equal(D, 2.5 * L.T)
    D, ierr = oracle.block_contract([], np.array(3.0), np.array(4.0), [])
    assert ierr == 0 and D[0] == 12.0


def test_contract_rejects_bad_extents(oracle):
    # contr_ptrn_ok (F90:861-896): extent mismatch -> ierr=1
    L = np.zeros((3, 4), order="F")
    R = np.zeros((5, 2), order="F")
    _, ierr = oracle.block_contract([1, -1, -2, 2], L, R, [3, 2])
    assert ierr == 1


def test_all_sial_patterns_vs_einsum(oracle):
    """Every distinct contraction pattern of the reference's CC / EOM SIAL programs (170, tests/golden): the oracle's
    pattern analysis + permute/dgemm/permute against numpy.einsum, non-uniform extents per index kind."""
    from conftest import sial_patterns

    pats = sial_patterns()
    assert len(pats) >= 150
    size = {"o": 3, "v": 5, "p": 4, "n": 6, "x": 2, "s": 2}
    rng = np.random.default_rng(42)
    for d, l, r, kinds, where in pats:
        num = {c: i + 1 for i, c in enumerate(kinds)}
        L = np.asfortranarray(rng.uniform(-1, 1, [size[kinds[c]] for c in l]))
        R = np.asfortranarray(rng.uniform(-1, 1, [size[kinds[c]] for c in r]))
        D, ierr = oracle.contract_labels([num[c] for c in d], [size[kinds[c]] for c in d], [num[c] for c in l], L,
                                         [num[c] for c in r], R)
        assert ierr == 0, where
        ref = np.einsum(f"{l},{r}->{d}", L, R)
        assert np.max(np.abs(D - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), where
