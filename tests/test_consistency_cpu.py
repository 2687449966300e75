"""Race detection between barriers: the library's order-independent section rule (sipgpu_consistency_validate, dist.cu)
against the oracle's row-by-row restatement of the reference's state table
(distributed_block_consistency.cpp:25-175) on exhaustive short and random long access sequences.  Host-only."""
import itertools
import random

import pytest


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


BITS = {0: 1, 1: 2, 2: 4}   # GET, PUT, PUT_ACCUMULATE -> SIPGPU_ACCESS_*


def library_says_ok(sip, ops, workers):
    try:
        sip.consistency_validate([(7, BITS[o], w) for o, w in zip(ops, workers)])
        return True
    except sip.SipGpuError:
        return False


def test_every_sequence_of_up_to_four_accesses(sip, oracle):
    for n in range(1, 5):
        for ops in itertools.product(range(3), repeat=n):
            for workers in itertools.product(range(3), repeat=n):
                want = oracle.block_consistency(list(ops), list(workers), [1] * n) == -1
                assert library_says_ok(sip, ops, workers) == want, (ops, workers)


def test_random_long_sequences_and_section_reset(sip, oracle):
    rnd = random.Random(4)
    for _ in range(300):
        n = rnd.randrange(5, 40)
        kind = rnd.randrange(4)
        ops = [rnd.randrange(3) if kind == 0 else (0 if kind == 1 else 2 if kind == 2 else rnd.choice([0, 2])) for _ in range(n)]
        workers = [rnd.randrange(4) if rnd.random() < 0.7 else 0 for _ in range(n)]
        want = oracle.block_consistency(ops, workers, [1] * n) == -1
        assert library_says_ok(sip, ops, workers) == want
    # a barrier in between makes put-by-one / get-by-another legal (sections are validated separately)
    assert oracle.block_consistency([1, 0], [0, 1], [1, 1]) == 1
    assert oracle.block_consistency([1, 0], [0, 1], [1, 2]) == -1
    assert library_says_ok(sip, [1], [0]) and library_says_ok(sip, [0], [1])


def test_known_reference_cases(sip, oracle):
    """Sial.put_accumulate_stress / get_mpi patterns (test_sial.cpp:485,583,1072): many accumulators, many readers"""
    assert library_says_ok(sip, [2] * 8, list(range(8)))            # put += from every worker
    assert library_says_ok(sip, [0] * 8, list(range(8)))            # get by every worker
    assert not library_says_ok(sip, [2, 0], [0, 1])                 # accumulate and read by another worker
    assert not library_says_ok(sip, [1, 1], [0, 1])                 # two writers
    assert library_says_ok(sip, [1, 0, 2, 1], [3, 3, 3, 3])         # a single worker may do anything
