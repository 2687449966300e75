import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


def sial_patterns():
    """The distinct contraction label patterns of the reference's SIAL programs (tests/golden/
    sial_contraction_patterns.txt, produced by scripts/extract_sial_patterns.py from src/sialx/qm):
    list of (dlabels, llabels, rlabels, kinds-by-label dict, where)."""
    out = []
    path = os.path.join(ROOT, "tests", "golden", "sial_contraction_patterns.txt")
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        d, l, r, kinds, count, where = line.split()
        labs = []
        for c in d + l + r:
            if c not in labs:
                labs.append(c)
        out.append((d, l, r, dict(zip(labs, kinds)), where))
    return out
