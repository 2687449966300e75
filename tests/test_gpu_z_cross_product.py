"""Config 4 of BASELINE.json, correctness leg ("run the full cross-product at s = 16 and
32 for correctness"): all 1728 rank-4 label patterns (tests/cross_product.py) through the fused CUDA contraction.
  * s = 16: every pattern against the CPU oracle (permute -> dgemm -> permute with OpenBLAS), 1e-10;
  * s = 32 (2.1 GFLOP per pattern, too slow for the oracle 1728 times): every pattern against the SAME contraction
    computed on the device in canonical label order and permuted with the (separately verified) permute kernel --
    a size-independent equivariance property -- with the canonical product itself checked against the oracle.
(tests/test_gpu_parity.py::test_sweep_rank4_full_cross_product_sample keeps the 24-pattern sample.  Written after the
round's GPU minutes were spent, hence the file name that sorts last.)"""
import numpy as np
import pytest

from cross_product import patterns
from test_gpu_parity import TOL, rand_block, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


@pytest.mark.timeout(600, method="thread")   # first GPU run pending: never hang the box
def test_full_cross_product_s16_against_the_oracle(sip, oracle, s=16):
    rng = np.random.default_rng(16)
    L0, R0 = rand_block(rng, (s,) * 4), rand_block(rng, (s,) * 4)
    dL, dR = sip.DeviceBlock.from_numpy(L0), sip.DeviceBlock.from_numpy(R0)
    out = sip.DeviceBlock((s,) * 4)
    oracle.use_openblas(8)
    worst = 0.0
    try:
        for dlab, llab, rlab in patterns():
            ref, oerr = oracle.contract_labels(dlab, [s] * 4, llab, L0, rlab, R0)
            assert oerr == 0
            sip.contract_labels(dlab, [s] * 4, llab, dL, rlab, dR, out=out)
            e = relerr(out.to_numpy().reshape(ref.shape), ref)
            assert e <= TOL, (dlab, llab, rlab, e)
            worst = max(worst, e)
    finally:
        oracle.use_naive_gemm()
    print(f"\n1728 patterns at s={s}: worst relative error {worst:.2e}")


@pytest.mark.timeout(600, method="thread")   # first GPU run pending: never hang the box
def test_full_cross_product_s32_equivariance(sip, oracle, s=32):
    rng = np.random.default_rng(32)
    L0, R0 = rand_block(rng, (s,) * 4), rand_block(rng, (s,) * 4)
    dL, dR = sip.DeviceBlock.from_numpy(L0), sip.DeviceBlock.from_numpy(R0)
    Lp, Rp, Dc, Dref, out = (sip.DeviceBlock((s,) * 4) for _ in range(5))
    canon_l, canon_r, canon_d = [5, 6, 1, 2], [5, 6, 3, 4], [1, 2, 3, 4]
    checked_canonical = set()
    worst = 0.0
    for n, (dlab, llab, rlab) in enumerate(patterns()):
        # reference on the device: operands permuted into canonical order, canonical contraction, result permuted
        sip.permute_labels(canon_l, llab, dL, out=Lp)      # Lp[5,6,1,2] = L[llab]
        sip.permute_labels(canon_r, rlab, dR, out=Rp)
        sip.contract_labels(canon_d, [s] * 4, canon_l, Lp, canon_r, Rp, out=Dc)
        sip.permute_labels(dlab, canon_d, Dc, out=Dref)    # Dref[dlab] = Dc[1,2,3,4]
        sip.contract_labels(dlab, [s] * 4, llab, dL, rlab, dR, out=out)
        ref = Dref.to_numpy()
        e = relerr(out.to_numpy(), ref)
        assert e <= TOL, (dlab, llab, rlab, e)
        worst = max(worst, e)
        # the canonical product depends on (llab, rlab) only through the permuted operands: check it against the oracle
        # once per operand arrangement class that changes the data (every 24th pattern starts a new (llab, rlab) pair);
        # 6 of the 72 pairs are enough to pin the device-side reference without an hour of CPU dgemm
        key = (tuple(llab), tuple(rlab))
        if n % (24 * 12) == 0 and key not in checked_canonical:
            checked_canonical.add(key)
            oracle.use_openblas(8)
            try:
                want, oerr = oracle.contract_labels(dlab, [s] * 4, llab, L0, rlab, R0)
            finally:
                oracle.use_naive_gemm()
            assert oerr == 0 and relerr(ref.reshape(want.shape), want) <= TOL
    assert len(checked_canonical) == 6
    print(f"\n1728 patterns at s={s}: worst relative difference to the canonical-order product {worst:.2e}")
