"""Host/device coherence of a block: the library's pure lazy_gpu_* transition (mirror.cu) against the oracle's
branch-by-branch restatement of BlockManager::lazy_gpu_{read,write,update}_on_{device,host}
(block_manager.cpp:340-441) for every status word and every operation, and along random operation sequences."""
import random

import pytest


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


def test_all_states_all_operations(sip, oracle):
    for bits in range(16):
        for op in range(6):
            rc, nb, act = sip.mirror_transition(bits, op)
            failed, onb, oact = oracle.lazy_gpu_transition(bits, op)
            assert (rc != 0) == (failed != 0), (bits, op)
            if not failed:
                assert (nb, act) == (onb, oact), (bits, op)
            else:
                assert rc == 105     # SIPGPU_E_STATE where the reference calls fail()


def test_random_walks_from_a_host_block(sip, oracle):
    rnd = random.Random(2)
    for _ in range(200):
        a = b = sip.ON_HOST
        for _ in range(30):
            op = rnd.randrange(6)
            rc, a2, act = sip.mirror_transition(a, op)
            failed, b2, oact = oracle.lazy_gpu_transition(b, op)
            assert rc == 0 and not failed and (a2, act) == (b2, oact)
            a, b = a2, b2
            assert a & (sip.ON_HOST | sip.ON_GPU)


def test_bad_arguments(sip):
    assert sip.mirror_transition(16, 0)[0] == 102 and sip.mirror_transition(1, 6)[0] == 102
