"""Product `_gpu_contract` / `_gpu_permute` (boundary 2, include/sipgpu.h) against the REFERENCE's own `_gpu_contract` /
`_gpu_permute`: src/sip/cuda/gpu_super_instructions.cu compiled unmodified for sm_100a (oracle/_ref/libaces4_ref_gpu.so,
`make -C oracle ref_gpu`) and run on the same B200 in a child process (oracle/ref_gpu.py).  This is the one parity check
of the contraction and the permutation whose other side is reference CODE rather than a restatement of it.

Cases: every contraction label pattern of the reference's SIAL programs (tests/golden/sial_contraction_patterns.txt) that
the legacy backend can express (at least one contracted index, block result, rank <= 6) at the block shapes of the
shipped test inputs (occ 5, virt 8, ao 13: lccd_test.dat) and one at the CCSD bench shape (50 x 20 x 50 x 20); the
known-answer transposes of test_basic_sial.cpp; all 24 rank-4 permutations.  Tolerance 1e-10 relative (north_star).

If the prebuilt reference library is absent (it is git-ignored; it travels with gpurun snapshots) or its child process
dies (the reference exits on any CUDA error), the test SKIPS with the reason -- it never passes vacuously.
"""
import ctypes as C
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


@pytest.fixture(scope="module")
def ref_gpu():
    from oracle import ref_gpu as r

    if not r.available():
        pytest.skip("oracle/_ref/libaces4_ref_gpu.so not built and no reference checkout here")
    return r


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def ia(v):
    v = [int(x) for x in v] + [1] * (6 - len(v))
    return (C.c_int * len(v))(*v)


def product_case(sip, ref_gpu, case):
    """the same case through libsipgpu's `_gpu_*` entry points (label arguments, device pointers)"""
    L = sip.lib()
    x1, x2 = ref_gpu.case_inputs(case)
    ydims, yinds = case["y"]
    y = np.zeros(tuple(ydims), order="F")
    g1 = L._gpu_allocate(x1.size)
    assert L._gpu_host_to_device(x1.ctypes.data_as(C.c_void_p), g1, x1.size) == 0
    gy = L._gpu_allocate(y.size)
    g2 = None
    if case["kind"] == "contract":
        g2 = L._gpu_allocate(x2.size)
        assert L._gpu_host_to_device(x2.ctypes.data_as(C.c_void_p), g2, x2.size) == 0
        rc = L._gpu_contract(gy, len(ydims), ia(ydims), ia(yinds), g1, x1.ndim, ia(case["x1"][0]), ia(case["x1"][1]),
                             g2, x2.ndim, ia(case["x2"][0]), ia(case["x2"][1]))
    else:
        rc = L._gpu_permute(gy, len(ydims), ia(ydims), ia(yinds), g1, x1.ndim, ia(case["x1"][0]), ia(case["x1"][1]))
    assert rc == 0, sip.lib().sipgpu_last_error()
    assert L._gpu_device_to_host(y.ctypes.data_as(C.c_void_p), gy, y.size) == 0
    for g in (g1, g2, gy):
        if g is not None:
            assert L._gpu_free(g) == 0
    return y


def contraction_cases():
    from conftest import sial_patterns

    sizes = dict(o=5, v=8, p=13, n=13, x=3, s=2)        # lccd_test.dat: occ [5], virt [8], ao [13]
    cases = []
    for k, (d, l, r, kinds, where) in enumerate(sial_patterns()):
        nc = len([c for c in l if c in r])
        if nc == 0 or len(d) == 0 or max(len(d), len(l), len(r)) > 6 or len(d) != len(l) + len(r) - 2 * nc:
            continue
        if len(set(l)) != len(l) or len(set(r)) != len(r):
            continue
        num = {c: i + 1 for i, c in enumerate(kinds)}
        ext = {c: sizes[kd] for c, kd in kinds.items()}
        cases.append({"kind": "contract", "seed": 1000 + k, "where": where,
                      "y": ([ext[c] for c in d], [num[c] for c in d]),
                      "x1": ([ext[c] for c in l], [num[c] for c in l]),
                      "x2": ([ext[c] for c in r], [num[c] for c in r])})
    # the CCSD bench block shape: ring term Z[a,j,b,i] = T[a,i,c,k] * V[b,j,c,k] on 50 x 20 x 50 x 20 blocks
    cases.append({"kind": "contract", "seed": 7, "where": "bench shape",
                  "y": ([50, 20, 50, 20], [1, 4, 3, 2]), "x1": ([50, 20, 50, 20], [1, 2, 5, 6]),
                  "x2": ([50, 20, 50, 20], [3, 4, 5, 6])})
    return cases


def permute_cases():
    cases = [  # BasicSial.transpose_tmp / transpose4d_tmp / transpose4d_square_tmp (test_basic_sial.cpp:653-693,1285-1406)
        {"kind": "permute", "seed": 1, "y": ([8, 8, 8], [2, 3, 1]), "x1": ([8, 8, 8], [1, 2, 3])},
        {"kind": "permute", "seed": 2, "y": ([5, 5, 5, 1], [3, 2, 1, 4]), "x1": ([5, 5, 5, 1], [1, 2, 3, 4])},
        {"kind": "permute", "seed": 3, "y": ([8, 8, 8, 8], [3, 2, 1, 4]), "x1": ([8, 8, 8, 8], [1, 2, 3, 4])},
    ]
    ext = {1: 8, 2: 5, 3: 9, 4: 6}
    for k, p in enumerate(itertools.permutations([1, 2, 3, 4])):
        cases.append({"kind": "permute", "seed": 100 + k, "y": ([ext[c] for c in p], list(p)),
                      "x1": ([ext[c] for c in (1, 2, 3, 4)], [1, 2, 3, 4])})
    return cases


def config4_cases():
    """BASELINE.json config 4 shapes the legacy backend can hold (blocks <= 40 MB: segment sizes 16..40): rank-4 results with
    two contracted indices under shuffled label orders, the ragged case (13, 30, 50, 64) and EOM-style rank-5 blocks with a
    leading extent-1 index"""
    import random

    pyrng = random.Random(2024)
    cases = []
    for s in (16, 24, 32, 40):
        for trial in range(4):
            fl, fr, cc = [1, 2], [3, 4], [5, 6]
            llab, rlab, dlab = fl + cc, fr + cc, fl + fr
            pyrng.shuffle(llab), pyrng.shuffle(rlab), pyrng.shuffle(dlab)
            cases.append({"kind": "contract", "seed": 3000 + 10 * s + trial, "where": f"s={s}",
                          "y": ([s] * 4, dlab), "x1": ([s] * 4, llab), "x2": ([s] * 4, rlab)})
    ext = {1: 13, 2: 30, 3: 50, 4: 64, 5: 9, 6: 11}
    for trial, (dlab, llab, rlab) in enumerate((([1, 2, 3, 4], [1, 5, 2, 6], [6, 3, 5, 4]), ([4, 3, 2, 1], [5, 1, 6, 2], [3, 5, 4, 6]),
                                                ([2, 4, 1, 3], [6, 5, 2, 1], [4, 6, 3, 5]))):
        cases.append({"kind": "contract", "seed": 3500 + trial, "where": "ragged",
                      "y": ([ext[c] for c in dlab], dlab), "x1": ([ext[c] for c in llab], llab), "x2": ([ext[c] for c in rlab], rlab)})
    ext = {1: 8, 2: 5, 3: 8, 4: 5, 5: 8, 6: 5, 7: 1}
    cases.append({"kind": "contract", "seed": 3600, "where": "EOM rank 5", "y": ([ext[c] for c in (7, 1, 2, 3, 4)], [7, 1, 2, 3, 4]),
                  "x1": ([ext[c] for c in (7, 1, 5, 2, 6)], [7, 1, 5, 2, 6]), "x2": ([ext[c] for c in (6, 3, 5, 4)], [6, 3, 5, 4])})
    return cases


def slab_cases():
    """rank-2 results of rank-4 blocks with a long contracted range at CCSD segment sizes (occupied 20, virtual 50, 8 MB operands):
    the shapes the TMA-fed slab kernel takes (lowint.cu: both operands in contiguous runs; one operand gathered; K order of the
    larger operand for the matrix-vector shapes) -- the reference's legacy backend permutes both operands and calls cuBLAS"""
    o, v = 20, 50
    cases = []
    for k, (d, l, r, ext) in enumerate((
            ("ab", "cade", "cbde", dict(a=o, b=o, c=v, d=v, e=o)),      # K-fast operands, 20 x 20
            ("ab", "acde", "bcde", dict(a=v, b=v, c=o, d=v, e=o)),      # M-fast operands, 50 x 50
            ("ab", "acde", "cbed", dict(a=o, b=v, c=o, d=v, e=o)),      # outer contracted indices in different orders
            ("ab", "acde", "cdeb", dict(a=o, b=v, c=v, d=v, e=o)),      # hybrid: one operand gathered
            ("ab", "acde", "bedc", dict(a=v, b=v, c=o, d=v, e=o)),      # hybrid: 400-byte pieces
            ("ab", "cdba", "dc", dict(a=o, b=v, c=v, d=v)))):           # matrix-vector, big operand K-fast
        labs = sorted(set(d + l + r))
        num = {c: i + 1 for i, c in enumerate(labs)}
        cases.append({"kind": "contract", "seed": 4000 + k, "where": f"slab {d}={l}*{r}",
                      "y": ([ext[c] for c in d], [num[c] for c in d]), "x1": ([ext[c] for c in l], [num[c] for c in l]),
                      "x2": ([ext[c] for c in r], [num[c] for c in r])})
    return cases


def run_reference(ref_gpu, cases):
    try:
        return ref_gpu.run_cases(cases, timeout=240)[0]
    except ref_gpu.WorkerFailed as e:
        # the fixture has already skipped when the prerequisite (prebuilt library / reference sources) is missing: a child
        # that crashes, times out or exits non-zero is a failure of the comparison, not a reason to skip it
        pytest.fail(f"reference CUDA backend child process failed: {e}")


def test_contractions_equal_the_reference_cuda_backend(sip, ref_gpu, oracle):
    cases = contraction_cases()
    assert len(cases) >= 100
    want = run_reference(ref_gpu, cases)
    worst = 0.0
    for case, w in zip(cases, want):
        got = product_case(sip, ref_gpu, case)
        err = relerr(got, w)
        worst = max(worst, err)
        assert err <= TOL, (case["where"], case["y"], case["x1"], case["x2"], err)
    # three-way on a sample: the oracle agrees with both
    for case, w in list(zip(cases, want))[::17]:
        x1, x2 = ref_gpu.case_inputs(case)
        o, ierr = oracle.contract_labels(case["y"][1], case["y"][0], case["x1"][1], x1, case["x2"][1], x2)
        assert ierr == 0 and relerr(o.reshape(w.shape), w) <= TOL
    print(f"{len(cases)} contraction patterns: product vs reference CUDA backend, worst relative error {worst:.2e}")


def test_permutations_equal_the_reference_cuda_backend(sip, ref_gpu, oracle):
    cases = permute_cases()
    want = run_reference(ref_gpu, cases)
    for case, w in zip(cases, want):
        got = product_case(sip, ref_gpu, case)
        assert np.array_equal(got, w), (case["y"], case["x1"])        # a permutation moves bits: exact
        x1, _ = ref_gpu.case_inputs(case)
        assert np.array_equal(oracle.permute_labels(case["y"][1], case["x1"][1], x1), w)


def test_config4_shapes_equal_the_reference_cuda_backend(sip, ref_gpu):
    cases = config4_cases()
    want = run_reference(ref_gpu, cases)
    for case, w in zip(cases, want):
        err = relerr(product_case(sip, ref_gpu, case), w)
        assert err <= TOL, (case["where"], case["y"], case["x1"], case["x2"], err)


def test_slab_shapes_equal_the_reference_cuda_backend(sip, ref_gpu):
    cases = slab_cases()
    want = run_reference(ref_gpu, cases)
    for case, w in zip(cases, want):
        err = relerr(product_case(sip, ref_gpu, case), w)
        assert err <= TOL, (case["where"], err)
