"""Deterministic case generators shared by scripts/make_ref_golden.py (which runs them through the reference's own C++,
oracle/_ref, and commits the answers as tests/golden/ref_vectors.json) and tests/test_oracle_vs_ref_cpu.py (which runs
them through the oracle, the product's host logic and -- where oracle/_ref is present -- the live reference again)."""
import hashlib
import itertools
import random

import numpy as np

CONSISTENCY_MAX_LEN = 5
CONSISTENCY_WORKERS = 3
VERDICT_CHARS = "-0123456789"   # '-' = accepted, digit = index of the first rejected access


def consistency_exhaustive(n):
    """every sequence of n accesses (op in GET/PUT/PUT_ACCUMULATE) by 3 workers inside one barrier section"""
    for ops in itertools.product(range(3), repeat=n):
        for workers in itertools.product(range(CONSISTENCY_WORKERS), repeat=n):
            yield list(ops), list(workers), [1] * n


def consistency_sectioned(count=400, seed=11):
    """random sequences with barrier sections (non-decreasing section numbers)"""
    rnd = random.Random(seed)
    for _ in range(count):
        n = rnd.randrange(2, 24)
        style = rnd.randrange(4)
        ops = [rnd.randrange(3) if style == 0 else (0 if style == 1 else 2 if style == 2 else rnd.choice([0, 2])) for _ in range(n)]
        workers = [rnd.randrange(4) if rnd.random() < 0.6 else 0 for _ in range(n)]
        sec, sections = 1, []
        for _ in range(n):
            if rnd.random() < 0.25:
                sec += rnd.randrange(1, 3)
            sections.append(sec)
        yield ops, workers, sections


def block_number_cases(seed=5):
    """(segment counts, lower segment values, index values): the arrays of the synthetic CCSD workload, the segment tables of
    the shipped .dat files, and random ranks 1..6"""
    rnd = random.Random(seed)
    fixed = [([12, 3, 12, 3], [1, 1, 1, 1]), ([3, 3, 3, 3], [1, 1, 1, 1]), ([1, 1, 1, 1], [2, 1, 2, 1]),
             ([2, 1, 2, 1], [3, 2, 3, 2]), ([4, 2], [1, 5])]
    for nseg, lower in fixed:
        for idx in itertools.product(*[range(lo, lo + n) for n, lo in zip(nseg, lower)]):
            if len(nseg) < 4 or rnd.random() < 0.08:
                yield nseg, lower, list(idx)
    for _ in range(300):
        rank = rnd.randrange(1, 7)
        nseg = [rnd.randrange(1, 7) for _ in range(rank)]
        lower = [rnd.randrange(1, 5) for _ in range(rank)]
        yield nseg, lower, [lo + rnd.randrange(n) for n, lo in zip(nseg, lower)]


def block_id_cases(seed=9):
    rnd = random.Random(seed)
    for _ in range(300):
        a = [rnd.randrange(1, 4) for _ in range(4)] + [-1, -1]    # unused_index_value pads the key
        b = list(a) if rnd.random() < 0.2 else [rnd.randrange(1, 4) for _ in range(4)] + [-1, -1]
        yield rnd.randrange(2), a, rnd.randrange(2), b


def transpose_cases():
    """(extents, 0-based permute vector: permute[i] = destination position of source dimension i)"""
    yield [8, 8, 8], [2, 0, 1]            # transpose_tmp: b[j,k,i] = a[i,j,k]
    yield [5, 5, 5, 1], [2, 1, 0, 3]      # transpose4d_tmp
    yield [8, 8, 8, 8], [2, 1, 0, 3]      # transpose4d_square_tmp
    for rank, ext in ((2, [7, 5]), (3, [4, 6, 3]), (4, [5, 3, 4, 2]), (5, [3, 2, 4, 2, 3]), (6, [2, 3, 2, 2, 3, 2])):
        perms = list(itertools.permutations(range(rank)))
        step = max(1, len(perms) // 24)
        for p in perms[::step]:
            yield ext, list(p)


def slice_cases(seed=3):
    rnd = random.Random(seed)
    for rank in range(1, 7):
        for _ in range(6):
            t_ext = [rnd.randrange(2, 7) for _ in range(rank)]
            s_ext = [rnd.randrange(1, e + 1) for e in t_ext]
            off = [rnd.randrange(0, e - s + 1) for e, s in zip(t_ext, s_ext)]
            yield t_ext, s_ext, off


def elementwise_cases():
    """(op name, extents, scalar)"""
    for ext in ([20, 20], [5, 8, 5, 8], [13], [3, 2, 4, 2, 3, 2]):
        for op, x in (("fill", 42.0), ("scale", -0.75), ("scale_and_copy", 1.0 / 3.0), ("copy_data", 0.0),
                      ("increment", 1e-3), ("accumulate", 0.0)):
            yield op, ext, x


def seeded(shape, tag):
    """inputs: splitmix-free, numpy-only so that the generator does not depend on anything under test"""
    rng = np.random.default_rng(0xACE54 + tag)
    return np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape))


def digest(a):
    return hashlib.sha256(np.asfortranarray(a, dtype=np.float64).tobytes(order="F")).hexdigest()[:24]


CHECKPOINT_SCALARS = {"scf_energy": -75.58432674274046, "lccd_correlation": -0.12610179886435, "padded label  ": 3.5}
CHECKPOINT_ARRAYS = [("ca", [3, 2], [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]), ("fock_a", [2, 2], [-20.25, 0.5, 0.5, -1.125])]
