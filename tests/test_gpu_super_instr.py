"""GPU parity of the on-device CC super-instructions (aces4_b200/csrc/superinstr.cu) against the oracle
(oracle/super_instr_oracle.c), through the C ABI with the reference's super-instruction calling convention.
Bit-exact: these are elementwise divisions / moves, evaluated in the reference's operation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEGS = [20, 16, 50, 34, 7]


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    s.api.set_predefined_int_array("moa_seg_ranges", SEGS)
    return s.api


def fblock(rng, shape):
    return np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape))


@pytest.mark.parametrize("fock_rank", [1, 2])
@pytest.mark.parametrize("iv", [(3, 1), (3, 1, 4, 2), (5, 5, 1, 2)])
def test_energy_denominator_rhf(sip, oracle, fock_rank, iv):
    rng = np.random.default_rng(len(iv) * 10 + fock_rank)
    n = sum(SEGS)
    diag = np.sort(rng.uniform(-20.0, 5.0, n)) + 0.37 * np.arange(n)
    fock = diag.copy() if fock_rank == 1 else np.asfortranarray(np.diag(diag) + 1e-3 * rng.uniform(-1, 1, (n, n)))
    blk = fblock(rng, [SEGS[v - 1] for v in iv])
    d = sip.DeviceBlock.from_numpy(blk)
    assert sip.si_energy_denominator_rhf(d, iv, sip.DeviceBlock.from_numpy(fock)) == 0
    assert oracle.si_energy_denominator_rhf(blk, iv, fock, SEGS) == 0
    assert np.array_equal(d.to_numpy(), blk)


def test_energy_denominator_rank6(sip, oracle):
    rng = np.random.default_rng(66)
    n = sum(SEGS)
    fock = np.asfortranarray(np.diag(np.sort(rng.uniform(-20.0, 5.0, n)) + 0.5 * np.arange(n)))
    iv = (2, 5, 2, 5, 30, 41)  # Xaaaiii[a,a1,b,k1,ii,jj]-like: four segment indices + two simple indices
    shape = [16, 7, 16, 7, 1, 1]
    blk = fblock(rng, shape)
    d = sip.DeviceBlock.from_numpy(blk)
    assert sip.si_energy_denominator_rhf(d, iv, sip.DeviceBlock.from_numpy(fock)) == 0
    assert oracle.si_energy_denominator_rhf(blk, iv, fock, SEGS) == 0
    assert np.array_equal(d.to_numpy(), blk)
    # unsupported rank and a block outside the Fock range are errors, not crashes
    assert sip.si_energy_denominator_rhf(sip.DeviceBlock((2, 2, 2)), (1, 1, 1), sip.DeviceBlock.from_numpy(fock)) != 0


def test_stripi(sip, oracle):
    rng = np.random.default_rng(8)
    iv0 = (3, 1, 3, 2)  # TSaiai[a2,i1,a,j1]
    x = fblock(rng, [SEGS[v - 1] for v in iv0])
    dx = sip.DeviceBlock.from_numpy(x)
    for jj in (21, 29, 36):  # global range of segment 2 is 21..36
        y = sip.DeviceBlock((50, 20, 50, 1), zero=True)
        assert sip.si_stripi(dx, iv0, y, (3, 1, 3, jj)) == 0
        ref, ierr = oracle.si_stripi(x, iv0, (50, 20, 50, 1), (3, 1, 3, jj), SEGS)
        assert ierr == 0
        assert np.array_equal(y.to_numpy(), ref)
    assert sip.si_stripi(dx, iv0, sip.DeviceBlock((50, 20, 50, 1)), (3, 1, 3, 37)) != 0  # outside: the reference aborts
    # rank 3 / rank 2 forms (rccsdpt_aab.sialx:557: stripi tppp[a,a2,i1] tpps[a,a2,ii])
    x3 = fblock(rng, (50, 50, 20))
    y3 = sip.DeviceBlock((50, 50, 1))
    assert sip.si_stripi(sip.DeviceBlock.from_numpy(x3), (3, 3, 1), y3, (3, 3, 7)) == 0
    assert np.array_equal(y3.to_numpy(), oracle.si_stripi(x3, (3, 3, 1), (50, 50, 1), (3, 3, 7), SEGS)[0])
    x2 = fblock(rng, (50, 20))
    y2 = sip.DeviceBlock((50, 1))
    assert sip.si_stripi(sip.DeviceBlock.from_numpy(x2), (3, 1), y2, (3, 20)) == 0
    assert np.array_equal(y2.to_numpy(), oracle.si_stripi(x2, (3, 1), (50, 1), (3, 20), SEGS)[0])


@pytest.mark.parametrize("iv", [(3, 1, 3, 1), (4, 2, 4, 2), (3, 1, 4, 1)])
def test_anti_symm_o_v(sip, oracle, iv):
    rng = np.random.default_rng(sum(iv))
    shape = [SEGS[v - 1] for v in iv]
    for name in ("si_anti_symm_o", "si_anti_symm_v"):
        x = fblock(rng, shape)
        d = sip.DeviceBlock.from_numpy(x)
        assert getattr(sip, name)(d, iv) == 0
        assert getattr(oracle, name)(x, iv, SEGS) == 0
        got = d.to_numpy()
        assert np.array_equal(got, x) and np.array_equal(np.signbit(got), np.signbit(x)), name
    assert sip.si_anti_symm_o(sip.DeviceBlock((4, 4)), (1, 1)) != 0


def test_anti_symm_v_simple_index_case(sip, oracle):
    segs = [1, 9]
    sip.set_predefined_int_array("moa_seg_ranges", segs)
    try:
        rng = np.random.default_rng(4)
        x = fblock(rng, (9, 1, 9, 1))
        d = sip.DeviceBlock.from_numpy(x)
        assert sip.si_anti_symm_v(d, (2, 1, 2, 1)) == 0
        assert oracle.si_anti_symm_v(x, (2, 1, 2, 1), segs) == 0
        got = d.to_numpy()
        assert np.array_equal(got, x) and np.array_equal(np.signbit(got), np.signbit(x))
    finally:
        sip.set_predefined_int_array("moa_seg_ranges", SEGS)


def test_return_sval_and_invert_diagonal(sip, oracle):
    rng = np.random.default_rng(12)
    s = sip.DeviceBlock((1,), zero=True)
    for shape in ((1, 1), (7,), (5, 3)):
        a = fblock(rng, shape)
        assert sip.si_return_sval(sip.DeviceBlock.from_numpy(a), s) == 0
        assert s.to_numpy()[0] == oracle.si_return_sval(a)[0]
    for shape in ((30, 20, 4), (3, 20, 7, 20, 5)):
        a1, a2 = fblock(rng, shape), fblock(rng, shape)
        a2.ravel(order="F")[::7] = 0.0
        d1 = sip.DeviceBlock.from_numpy(a1)
        assert sip.si_invert_diagonal(d1, sip.DeviceBlock.from_numpy(a2)) == 0
        assert oracle.si_invert_diagonal(a1, a2) == 0
        assert np.array_equal(d1.to_numpy(), a1)
    assert sip.si_invert_diagonal(sip.DeviceBlock((2, 2)), sip.DeviceBlock((2, 2))) != 0


def test_return_diagonal_elements_and_invert_diagonal_asym(sip, oracle):
    rng = np.random.default_rng(21)
    for shape, iv in (((8, 8), (2, 2)), ((5, 5, 8, 8), (1, 1, 2, 2)), ((20, 20, 50, 50), (1, 1, 2, 2))):
        x = fblock(rng, shape)
        d = sip.DeviceBlock.from_numpy(x)
        assert sip.si_return_diagonal_elements(d, iv) == 0
        assert oracle.si_return_diagonal_elements(x, iv, SEGS) == 0
        assert np.array_equal(d.to_numpy(), x)
    assert sip.si_return_diagonal_elements(sip.DeviceBlock((2, 3)), (1, 1)) != 0
    for shape, iv in (((1, 8, 5, 8, 5), (3, 2, 1, 2, 1)), ((1, 50, 20, 50, 20), (7, 2, 1, 2, 1)), ((2, 8, 5, 5, 8), (1, 2, 1, 1, 2))):
        a1, a2 = fblock(rng, shape), fblock(rng, shape)
        a2.ravel(order="F")[::9] = 0.0
        d1 = sip.DeviceBlock.from_numpy(a1)
        assert sip.si_invert_diagonal_asym(d1, iv, sip.DeviceBlock.from_numpy(a2)) == 0
        assert oracle.si_invert_diagonal_asym(a1, iv, a2, SEGS) == 0
        assert np.array_equal(d1.to_numpy(), a1), shape
    assert sip.si_invert_diagonal_asym(sip.DeviceBlock((2, 2, 2)), (1, 1, 1), sip.DeviceBlock((2, 2, 2))) != 0


def test_energy_ty_denominator_rhf(sip, oracle):
    rng = np.random.default_rng(33)
    n = sum(SEGS)
    fock = np.asfortranarray(np.diag(np.sort(rng.uniform(-3, 3, n))) + 0.01 * rng.uniform(-1, 1, (n, n)))
    dfock = sip.DeviceBlock.from_numpy(fock)
    for iv, shift in (((3, 1, 3, 1), 0.4342791), ((4, 2, 3, 1), -0.125), ((5, 1, 5, 2), 0.0)):
        shape = tuple(SEGS[s - 1] for s in iv)
        x = fblock(rng, shape)
        d = sip.DeviceBlock.from_numpy(x)
        assert sip.si_energy_ty_denominator_rhf(d, iv, dfock, sip.DeviceBlock((1,)).fill(shift)) == 0
        assert oracle.si_energy_ty_denominator_rhf(x, iv, fock, shift, SEGS) == 0
        assert np.array_equal(d.to_numpy(), x), iv
    assert sip.si_energy_ty_denominator_rhf(sip.DeviceBlock((4, 4)), (1, 1), dfock, sip.DeviceBlock((1,)).fill(0.0)) != 0
