"""Link-level drop-in check, no GPU needed: the reference's own block.cpp (compiled in place) links against libsipgpu.so
with nothing left undefined -- the product exports tensor_block_copy__/slice__/insert__ under the reference's names and
calling convention (tensor_ops_c_prototypes.h:41-178) -- and, on a machine without a GPU, a call fails the way the
reference fails for any backend error: non-zero `ierr` -> CHECK -> sip::fail (block.cpp:252), never a silent CPU result."""
import ctypes as C
import subprocess

import pytest


@pytest.fixture(scope="module")
def rb():
    import aces4_b200 as s

    s.build()
    from oracle import ref_on_sipgpu as r

    if not r.available():
        pytest.skip("no reference checkout and no prebuilt oracle/_ref on this machine")
    return r


def test_reference_block_objects_link_against_the_product(rb):
    so = rb.build()
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True, check=True).stdout
    for sym in ("tensor_block_copy__", "tensor_block_slice__", "tensor_block_insert__"):
        assert f"U {sym}" in undefined                      # imported, not defined by the test library itself
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True, check=True).stdout
    assert "libsipgpu.so" in needed and "liboracle" not in needed
    C.CDLL(so)                                              # every symbol resolves at load time (built with --no-undefined)


def test_without_a_gpu_the_reference_fails_loudly_through_its_own_check(rb):
    import aces4_b200 as s

    if s.api.lib().sipgpu_init(-1) == 0:
        pytest.skip("a GPU is present: the success path is tests/test_gpu_ref_block_on_sipgpu.py")
    with pytest.raises(rb.WorkerFailed) as e:
        rb.run([{"op": "transpose", "ext": [3, 4, 5], "permute": [2, 0, 1], "seed": 1}])
    assert "error returned from tensor_block_copy_" in str(e.value)


def test_level1_reference_block_gpu_methods_link_against_the_product(rb):
    """INTEGRATION.md level 1 at link level: block.cpp compiled with HAVE_CUDA against the replacement
    gpu_super_instructions.h (oracle/ref_shim/level1: three lines around sipgpu.h) imports the `_gpu_*` names unmangled
    and libsipgpu.so provides all of them."""
    rb.build()
    undefined = subprocess.run(["nm", "-D", "--undefined-only", rb._SO_L1], capture_output=True, text=True, check=True).stdout
    for sym in ("_gpu_allocate", "_gpu_free", "_gpu_double_memset", "_gpu_selfmultiply", "_gpu_device_to_device"):
        assert f"U {sym}\n" in undefined, sym
    C.CDLL(rb._SO_L1)


def test_level1_without_a_gpu_no_device_block_is_handed_out(rb):
    import aces4_b200 as s

    if s.api.lib().sipgpu_init(-1) == 0:
        pytest.skip("a GPU is present: the success path is tests/test_gpu_ref_block_on_sipgpu.py")
    with pytest.raises(rb.WorkerFailed) as e:
        rb.run([{"op": "gpu_block", "ext": [4, 3], "fill": 2.0, "scale": 0.5}])
    assert "rc=2" in str(e.value)


def test_level1_interpreter_compiles_against_the_replacement_header(rb):
    """interpreter.cpp (reference, unmodified) with HAVE_CUDA and the level-1 replacement of gpu_super_instructions.h:
    its `_init_gpu(&devid, &rank)` call (interpreter.cpp:84-88) type-checks against include/sipgpu.h"""
    import os

    src = os.path.join(rb.REFERENCE_ROOT, "src", "sip")
    if not os.path.isdir(src):
        pytest.skip("no reference checkout on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = ["-I" + os.path.join(root, "oracle", "ref_shim", "level1"), "-I" + os.path.join(root, "include")]
    inc += ["-I" + os.path.join(src, d) for d in ("core", "cuda", "dynamic_data", "mpi", "setup", "static_data",
                                                   "super_instructions", "tensor_algebra", "worker", ".")]
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-w", *inc, os.path.join(src, "worker", "interpreter.cpp")])


def test_level2_sial_ops_adapter_compiles_and_links_against_the_reference(rb):
    """INTEGRATION.md level 2 at link level: include/sial_ops_device_aces4.hpp -- SialOpsDevice behind the reference's
    SialOpsParallel signatures (sial_ops_parallel.h:47-73: BlockId&, Block::BlockPtr, pc) -- compiles against the reference's
    block_id.h / block.h (HAVE_CUDA) and links, --no-undefined, against the reference objects and libsipgpu.so."""
    rb.build()
    import os

    assert os.path.exists(rb._SO_L2)
    lib = C.CDLL(rb._SO_L2)
    assert hasattr(lib, "aces4ref_l2_selftest")
    needed = subprocess.run(["readelf", "-d", rb._SO_L2], capture_output=True, text=True, check=True).stdout
    assert "libsipgpu.so" in needed and "liboracle" not in needed


def test_level2_without_a_gpu_fails_loudly(rb):
    import aces4_b200 as s

    if s.api.lib().sipgpu_init(-1) == 0:
        pytest.skip("a GPU is present: the success path is tests/test_gpu_ref_block_on_sipgpu.py")
    with pytest.raises(rb.WorkerFailed) as e:
        rb.run_level2_selftest()
    assert "no CUDA device" in str(e.value)
