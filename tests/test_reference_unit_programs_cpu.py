"""The reference's block-operation unit tests driven from their OWN SIAL programs (tests/ref_unit_programs.py) on the oracle
backend.  GPU twin: tests/test_gpu_z_reference_unit_programs.py."""
import pytest

import ref_unit_programs as rp
from sial_oracle_backend import OracleBackend


def make_backend(prog, seg_tables, constants):
    from oracle import oracle
    arrays = {n: {} for n, (k, _) in prog.arrays.items() if k in ("served", "distributed")}
    return OracleBackend(oracle, arrays, moa_seg_ranges=seg_tables["mo"] or [1])


@pytest.mark.parametrize("case", rp.ALL, ids=lambda f: f.__name__)
def test_reference_unit_program(case):
    case(make_backend, lambda h: h.a)


@pytest.mark.parametrize("case", rp.HOST_ONLY, ids=lambda f: f.__name__)
def test_reference_host_only_unit_program(case):
    """the pardo work distribution (BalancedTaskAllocParallelPardoLoop::do_update) and the interpreter's scalar / int /
    if-else arithmetic through the reference's own test programs: no block operation, so no device twin"""
    case(make_backend, lambda h: h.a)
