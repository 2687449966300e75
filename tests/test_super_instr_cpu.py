"""CPU tests of the super-instruction oracle (oracle/super_instr_oracle.c) against independent numpy statements of
energy_denominator_rhf.F, stripi.F, anti_symm_o/v.F, return_sval.F, invert_diagonal.F, invert_diagonal_asym.F and
return_diagonal_elements.F."""
import numpy as np
import pytest

SEGS = [3, 4, 2, 5]  # moa_seg_ranges: extents of the MO segments; global offsets 0, 3, 7, 9


def offs(iv):
    return [sum(SEGS[: v - 1]) for v in iv]


@pytest.mark.parametrize("fock_rank", [1, 2])
@pytest.mark.parametrize("iv", [(1, 2), (2, 1, 3, 4), (4, 4, 1, 2)])
def test_energy_denominator(oracle, fock_rank, iv):
    rng = np.random.default_rng(len(iv) + fock_rank)
    n = sum(SEGS)
    diag = np.sort(rng.uniform(-2.0, 3.0, n)) + np.arange(n)  # distinct orbital energies
    fock = diag.copy() if fock_rank == 1 else np.asfortranarray(np.diag(diag) + 1e-3 * rng.uniform(-1, 1, (n, n)))
    if fock_rank == 2:
        diag = np.diag(fock).copy()
    shape = [SEGS[v - 1] for v in iv]
    blk = np.asfortranarray(rng.uniform(-1, 1, shape))
    ref = blk.copy()
    o = offs(iv)
    for idx in np.ndindex(*shape):
        e = [diag[idx[d] + o[d]] for d in range(len(iv))]
        eps = (e[1] - e[0]) if len(iv) == 2 else (e[1] + e[3] - e[0] - e[2])
        ref[idx] = ref[idx] / eps
    assert oracle.si_energy_denominator_rhf(blk, iv, fock, SEGS) == 0
    assert np.array_equal(blk, ref)


def test_energy_denominator_rank6_simple_indices(oracle):
    rng = np.random.default_rng(6)
    n = sum(SEGS)
    fock = np.asfortranarray(np.diag(np.arange(1.0, n + 1) * 1.5))
    iv = (2, 1, 3, 2, 5, 7)          # last two are simple indices (extent 1): offset = value - 1
    shape = [4, 3, 2, 4, 1, 1]
    blk = np.asfortranarray(rng.uniform(-1, 1, shape))
    ref = blk.copy()
    o = offs(iv[:4]) + [iv[4] - 1, iv[5] - 1]
    d = np.diag(fock)
    with np.errstate(divide="ignore"):   # this synthetic diagonal has vanishing denominators: +-inf, as the reference's loop gives
        for idx in np.ndindex(*shape):
            e = [d[idx[k] + o[k]] for k in range(6)]
            ref[idx] /= e[1] + e[3] + e[5] - e[0] - e[2] - e[4]
    assert not np.isfinite(ref).all() and np.isfinite(ref).any()
    assert oracle.si_energy_denominator_rhf(blk, iv, fock, SEGS) == 0
    assert np.array_equal(blk, ref)
    assert oracle.si_energy_denominator_rhf(np.zeros((2, 2, 2), order="F"), (1, 1, 1), fock, SEGS) == 1  # rank 3: unsupported


def test_stripi(oracle):
    rng = np.random.default_rng(1)
    # TSaiai[a2,i1,a,j1] (segments 2,1,4,2) -> tppps[a2,i1,a,jj] with the simple index jj = global occupied index
    iv0 = (2, 1, 4, 2)
    x = np.asfortranarray(rng.uniform(-1, 1, [SEGS[v - 1] for v in iv0]))
    for jj in (4, 5, 6, 7):  # global range of segment 2 is 4..7
        y, ierr = oracle.si_stripi(x, iv0, (4, 3, 5, 1), (2, 1, 4, jj), SEGS)
        assert ierr == 0
        assert np.array_equal(y[..., 0], x[..., jj - 4])
    assert oracle.si_stripi(x, iv0, (4, 3, 5, 1), (2, 1, 4, 8), SEGS)[1] == 2  # the reference aborts: index outside the block
    # rank 3 with TWO stripped dimensions: a size-1 segment in the middle + the simple index
    segs = [3, 1, 4]
    x3 = np.asfortranarray(rng.uniform(-1, 1, (3, 1, 4)))
    lib_y, ierr = oracle.si_stripi(x3, (1, 2, 3), (3, 1, 1), (1, 2, 6), segs)
    assert ierr == 0 and np.array_equal(lib_y[:, 0, 0], x3[:, 0, 6 - 5])


def test_anti_symm(oracle):
    rng = np.random.default_rng(2)
    segs = [4, 3]
    for iv in ((1, 2, 1, 2), (1, 1, 1, 1)):
        shape = [segs[v - 1] for v in iv]
        x0 = np.asfortranarray(rng.uniform(-1, 1, shape))
        xo = x0.copy(order="F")
        assert oracle.si_anti_symm_o(xo, iv, segs) == 0
        ref = x0.copy()
        for a, i, b, j in np.ndindex(*shape):
            if i < j:
                ref[a, j, b, i] = -x0[a, i, b, j]
        for a, i, b, j in np.ndindex(*shape):
            if i == j or a == b:
                ref[a, i, b, j] = 0.0
        assert np.array_equal(xo, ref)
        xv = x0.copy(order="F")
        assert oracle.si_anti_symm_v(xv, iv, segs) == 0
        ref = x0.copy()
        for a, i, b, j in np.ndindex(*shape):
            if a < b:
                ref[b, i, a, j] = -x0[a, i, b, j]
        for a, i, b, j in np.ndindex(*shape):
            if i == j or a == b:
                ref[a, i, b, j] = 0.0
        assert np.array_equal(xv, ref)
    assert oracle.si_anti_symm_o(np.zeros((2, 2), order="F"), (1, 1), segs) == 1


def test_anti_symm_v_simple_index_case(oracle):
    # i range == 1:1 (anti_symm_v.F special case): mirror and write -0.0 on the a == b diagonal, i == j is NOT zeroed
    segs = [1, 3]
    iv = (2, 1, 2, 1)
    x0 = np.asfortranarray(np.arange(1.0, 10.0).reshape(3, 1, 3, 1, order="F"))
    x = x0.copy(order="F")
    assert oracle.si_anti_symm_v(x, iv, segs) == 0
    for a in range(3):
        for b in range(3):
            if a < b:
                assert x[b, 0, a, 0] == -x0[a, 0, b, 0] and x[a, 0, b, 0] == x0[a, 0, b, 0]
        assert x[a, 0, a, 0] == 0.0 and np.signbit(x[a, 0, a, 0])


def test_return_sval_and_invert_diagonal(oracle):
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.uniform(-1, 1, (1, 1)))
    assert oracle.si_return_sval(a) == (a[0, 0], 0)
    v = np.asfortranarray(rng.uniform(-1, 1, (5,)))
    assert oracle.si_return_sval(v) == (v[4], 0)          # doreturn1: array1(a2)
    m = np.asfortranarray(rng.uniform(-1, 1, (3, 4)))
    assert oracle.si_return_sval(m) == (m[2, 3], 0)        # doreturn2: array1(a2,b2)
    for shape in ((3, 4, 2), (2, 3, 1, 2, 3)):
        a1 = np.asfortranarray(rng.uniform(-1, 1, shape))
        a2 = np.asfortranarray(rng.uniform(-1, 1, shape))
        a2[tuple(0 for _ in shape)] = 0.0
        ref = np.where(a2 != 0.0, a1 / np.where(a2 != 0.0, a2, 1.0), a1)
        assert oracle.si_invert_diagonal(a1, a2) == 0
        assert np.array_equal(a1, ref)
    assert oracle.si_invert_diagonal(np.zeros((2, 2), order="F"), np.zeros((2, 2), order="F")) == 1


def test_return_diagonal_elements_and_invert_diagonal_asym(oracle):
    """return_diagonal_elements.F: x(p,p) / x(p,p,r,r) survive, by POSITION inside the block (both dimensions of a pair are
    declared over the first one's range); invert_diagonal_asym.F: rank 5, a1 /= a2 where the orbital numbers b != d and
    c != e (segment offsets count) and a2 != 0, a1 = 0 on the b == d or c == e planes"""
    rng = np.random.default_rng(8)
    segs = [2, 3, 3, 5]
    x = np.asfortranarray(rng.uniform(-1, 1, (4, 4)))
    ref = np.diag(np.diag(x))
    assert oracle.si_return_diagonal_elements(x, (3, 4), segs) == 0 and np.array_equal(x, ref)
    x = np.asfortranarray(rng.uniform(-1, 1, (3, 3, 5, 5)))
    ref = np.zeros_like(x)
    for p in range(3):
        for r in range(5):
            ref[p, p, r, r] = x[p, p, r, r]
    assert oracle.si_return_diagonal_elements(x, (2, 3, 4, 4), segs) == 0 and np.array_equal(x, ref)
    assert oracle.si_return_diagonal_elements(np.zeros((2, 3), order="F"), (1, 2), segs) == 1
    assert oracle.si_return_diagonal_elements(np.zeros((2, 2, 2), order="F"), (1, 1, 1), segs) == 1
    # blocks [k, a, i, a1, i1]: a in virtual segment 3 (orbitals 6..8), a1 in segment 4 (9..13) or 3; i, i1 in segments 1 / 2
    off = np.concatenate([[0], np.cumsum(segs)])
    for iv in ((1, 3, 1, 3, 1), (1, 3, 1, 4, 2), (2, 4, 2, 3, 2)):
        shape = (1,) + tuple(segs[s - 1] for s in iv[1:])
        a1 = np.asfortranarray(rng.uniform(-1, 1, shape))
        a2 = np.asfortranarray(rng.uniform(-1, 1, shape))
        a2.ravel(order="F")[::5] = 0.0
        ref = a1.copy(order="F")
        for idx in np.ndindex(*shape):
            g = [idx[d] + off[iv[d] - 1] for d in range(1, 5)]
            if g[0] != g[2] and g[1] != g[3]:
                if a2[idx] != 0.0:
                    ref[idx] = a1[idx] / a2[idx]
            else:
                ref[idx] = 0.0
        assert oracle.si_invert_diagonal_asym(a1, iv, a2, segs) == 0
        assert np.array_equal(a1, ref), iv
    assert oracle.si_invert_diagonal_asym(np.zeros((2, 2, 2), order="F"), (1, 1, 1), np.zeros((2, 2, 2), order="F"), segs) == 1


def test_energy_ty_denominator_rhf(oracle):
    """energy_ty_denominator_rhf.F do_rhfty_den4: x(a,b,c,d) /= eps_b + eps_d - eps_a - eps_c + shift, orbital numbers from the
    segment offsets"""
    rng = np.random.default_rng(19)
    segs = [2, 3, 3, 5]
    off = np.concatenate([[0], np.cumsum(segs)])
    eps = np.sort(rng.uniform(-2, 2, 13))
    fock = np.asfortranarray(np.diag(eps) + 0.01 * rng.uniform(-1, 1, (13, 13)))
    for iv, shift in (((3, 1, 4, 2), 0.375), ((4, 2, 4, 2), -0.2)):
        shape = tuple(segs[s - 1] for s in iv)
        x = np.asfortranarray(rng.uniform(-1, 1, shape))
        ref = x.copy(order="F")
        for a, b, c, d in np.ndindex(*shape):
            e = [fock[i + off[s - 1], i + off[s - 1]] for i, s in zip((a, b, c, d), iv)]
            ref[a, b, c, d] = x[a, b, c, d] / (e[1] + e[3] - e[0] - e[2] + shift)
        assert oracle.si_energy_ty_denominator_rhf(x, iv, fock, shift, segs) == 0
        assert np.array_equal(x, ref)
    assert oracle.si_energy_ty_denominator_rhf(np.zeros((2, 2), order="F"), (1, 1), fock, 0.0, segs) == 1
