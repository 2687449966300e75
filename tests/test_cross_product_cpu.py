"""CPU side of the config-4 cross product (tests/cross_product.py): all 1728 label patterns through the PRODUCT's
planner entry point `get_contraction_ptrn_` (host code of libsipgpu.so, runs without a GPU) vs the oracle's
restatement of tensor_dil_omp.F90:87-142, and through the oracle's contraction vs numpy.einsum at small ragged
extents.  The CUDA kernel on the same 1728 patterns: tests/test_gpu_z_cross_product.py."""
import numpy as np

import aces4_b200
from cross_product import einsum_spec, patterns


def test_the_cross_product_has_1728_distinct_patterns():
    pats = [tuple(map(tuple, p)) for p in patterns()]
    assert len(pats) == 1728 and len(set(pats)) == 1728


def test_planner_patterns_match_the_oracle_on_the_full_cross_product(oracle):
    sip = aces4_b200.api
    for dlab, llab, rlab in patterns():
        got, ierr = sip.get_contraction_ptrn(dlab, llab, rlab)
        want, oerr = oracle.get_contraction_ptrn(dlab, llab, rlab)
        assert ierr == 0 and oerr == 0 and list(got) == list(want), (dlab, llab, rlab)


def test_oracle_contraction_matches_einsum_on_the_full_cross_product(oracle):
    rng = np.random.default_rng(4)
    ext = {1: 2, 2: 3, 3: 4, 4: 2, 5: 3, 6: 5}
    worst = 0.0
    for dlab, llab, rlab in patterns():
        L = np.asfortranarray(rng.uniform(-1, 1, [ext[x] for x in llab]))
        R = np.asfortranarray(rng.uniform(-1, 1, [ext[x] for x in rlab]))
        got, ierr = oracle.contract_labels(dlab, [ext[x] for x in dlab], llab, L, rlab, R)
        assert ierr == 0
        want = np.einsum(einsum_spec(dlab, llab, rlab), L, R)
        worst = max(worst, np.max(np.abs(got - want)) / np.max(np.abs(want)))
    assert worst <= 1e-13
