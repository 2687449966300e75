"""The reference's block-operation unit tests (test/test_basic_sial.cpp, test/test_sial.cpp) driven from their OWN SIAL programs
(tests/golden/ref_unit_programs/*.sialx = src/sialx/test/*.sialx verbatim) through the SIAL front-end on libsipgpu: every
contraction, transpose, block add / subtract / scale, scalar contraction, put / get / put += is a C-ABI call; the assertions are
the C++ tests' (tests/ref_unit_programs.py).  Each program runs recorded (one work-list per pardo) and op-at-a-time.
CPU twin: tests/test_reference_unit_programs_cpu.py."""
import pytest

import ref_unit_programs as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init()
    return s.api


def backend_factory(sip, record):
    from aces4_b200.sial_frontend import DeviceBackend

    def make_backend(prog, seg_tables, constants):
        sip.set_predefined_int_array("moa_seg_ranges", seg_tables["mo"] or [1])
        arrays = {}
        for name, (kind, decl) in prog.arrays.items():
            if kind in ("served", "distributed"):
                arrays[name] = sip.DistArray([rp.dim_segments(prog, d, seg_tables, constants) for d in decl])
                arrays[name].fill_local(0.0)
        return DeviceBackend(sip, arrays, record=record)

    return make_backend


@pytest.mark.parametrize("record", [True, False], ids=["recorded", "op_at_a_time"])
@pytest.mark.parametrize("case", rp.ALL, ids=lambda f: f.__name__)
def test_reference_unit_program_on_the_device(sip, case, record):
    l0 = sip.kernel_launches()
    case(backend_factory(sip, record), lambda h: h.to_numpy())
    assert sip.kernel_launches() > l0
