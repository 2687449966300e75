"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/sipgpu.h declares, its host-only
planner entry points agree with the oracle, and compute calls FAIL LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    s.lib()
    return s.api


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sipgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = text[text.index('extern "C" {'):]
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", body)
    return sorted({n for n in names if n.startswith(("tensor_", "get_contraction", "_gpu_", "_init_gpu", "_finalize_gpu",
                                                      "sipgpu_"))})


def test_every_declared_symbol_is_exported(sip):
    syms = header_symbols()
    assert len(syms) >= 60
    for group in (sip.BOUNDARY1, sip.BOUNDARY2):
        for name in group:
            assert name in syms
    L = sip.lib()
    for name in syms:
        assert hasattr(L, name), f"{name} declared in include/sipgpu.h but not exported by libsipgpu.so"


def test_nothing_but_the_declared_abi_is_exported(sip):
    """the dynamic symbol table is exactly include/sipgpu.h (aces4_b200/csrc/exports.map): no internal C++ symbol leaks"""
    import subprocess
    import aces4_b200

    out = subprocess.check_output(["nm", "-D", "--defined-only", aces4_b200.lib_path()]).decode()
    exported = {ln.split()[-1] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] in "TBD"}
    assert exported == set(header_symbols()), (sorted(exported - set(header_symbols())), sorted(set(header_symbols()) - exported))


def test_header_is_plain_c99_and_links_from_c(sip, tmp_path):
    """include/sipgpu.h is what a C or Fortran (iso_c_binding) caller sees: it must be valid C99 on its own, with and
    without the libtensordil prototypes, and a C program must link against the library"""
    import subprocess
    import aces4_b200

    src = tmp_path / "client.c"
    src.write_text('#include "sipgpu.h"\n#include <stdio.h>\nint main(void) { printf("%d\\n", sipgpu_device()); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-DSIPGPU_NO_TENSORDIL_PROTOTYPES",
                           "-I", inc, str(src)])
    lib_dir = os.path.dirname(aces4_b200.lib_path())
    exe = tmp_path / "client"
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-L", lib_dir, "-lsipgpu", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    assert subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip() in ("-1", "0")


def test_library_links_no_oracle():
    # the product must not link, load or call anything under oracle/
    import subprocess
    import aces4_b200

    out = subprocess.check_output(["ldd", aces4_b200.lib_path()]).decode()
    assert "oracle" not in out
    strings = subprocess.check_output(["nm", "-D", aces4_b200.lib_path()]).decode()
    assert "oracle_" not in strings
    for dirpath, _, files in os.walk(os.path.join(ROOT, "aces4_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_get_contraction_ptrn_matches_oracle(sip, oracle):
    # host-only planner (F90:87-142): identical patterns and error codes on random label triples
    rng = random.Random(5)
    assert sip.get_contraction_ptrn([1, 4], [1, 2, 3, 4], [2, 3]) == ([1, -1, -2, 2, -2, -3], 0)
    for _ in range(3000):
        nfl, nfr, nc = rng.randint(0, 3), rng.randint(0, 3), rng.randint(0, 3)
        labels = rng.sample(range(1, 40), nfl + nfr + nc)
        fl, fr, cc = labels[:nfl], labels[nfl:nfl + nfr], labels[nfl + nfr:]
        llab, rlab, dlab = fl + cc, fr + cc, fl + fr
        rng.shuffle(llab), rng.shuffle(rlab), rng.shuffle(dlab)
        if rng.random() < 0.3 and (llab or rlab or dlab):  # corrupt: duplicate / drop / rename a label
            tgt = rng.choice([x for x in (llab, rlab, dlab) if x] )
            op = rng.random()
            if op < 0.4:
                tgt[rng.randrange(len(tgt))] = rng.randint(1, 40)
            elif op < 0.7:
                tgt.pop()
            else:
                tgt.append(rng.choice(labels) if labels else 1)
        got = sip.get_contraction_ptrn(dlab, llab, rlab)
        ref = oracle.get_contraction_ptrn(dlab, llab, rlab)
        assert got[1] == ref[1], (dlab, llab, rlab)
        if ref[1] == 0:
            assert got[0] == ref[0], (dlab, llab, rlab)


def test_tensor_size_by_shape(sip, oracle):
    assert sip.tensor_size_by_shape([3, 4, 5]) == (60, 0)
    assert sip.tensor_size_by_shape([]) == (1, 0)
    assert sip.tensor_size_by_shape([3, 0])[1] == 2


def test_compute_fails_loudly_without_gpu(sip):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device path cannot be exercised")
    # boundary 1: ierr is set, output untouched
    a = np.ones((3, 4), order="F")
    b = np.ones((4, 5), order="F")
    d, ierr = sip.tensor_block_contract([1, -1, -1, 2], a, b, [3, 5])
    assert ierr == 100 and np.all(np.isnan(d))  # SIPGPU_E_NODEVICE
    out, ierr = sip.tensor_block_copy(a, [1, 2, 1])
    assert ierr == 100 and np.all(np.isnan(out))
    # boundary 2/3: status codes / NULL, and a readable message
    L = sip.lib()
    assert L._init_gpu(C.byref(C.c_int(0)), C.byref(C.c_int(0))) == 100
    assert not L._gpu_allocate(16)
    with pytest.raises(sip.SipGpuError, match="no CUDA device"):
        sip.init()
    with pytest.raises(sip.SipGpuError):
        sip.DeviceBlock((4, 4))


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    import aces4_b200.api as api

    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(api, "_HERE", str(tmp_path))
    with pytest.raises(api.SipGpuError, match="no CPU fallback"):
        api.lib()


def test_trace_table_counts_calls_per_entry_point(sip):
    """sipgpu_trace_*: the per-entry-point table (reference: Tracer's per-opcode histogram and timer, tracer.h:41-50) --
    host-only entry points are enough to see it count"""
    api = sip.api if hasattr(sip, "api") else sip
    api.trace(True)
    for _ in range(5):
        api.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    api.lib().sipgpu_set_tuning(b"lowint_scope", 1.0)
    rows = {name: (calls, secs) for name, calls, secs in api.trace_report()}
    api.trace(False, reset=False)
    assert rows["get_contraction_ptrn_"][0] == 5 and rows["get_contraction_ptrn_"][1] >= 0.0
    assert rows["sipgpu_set_tuning"][0] == 1
    api.get_contraction_ptrn([1, 2], [1, 3], [3, 2])      # tracing off: not counted
    assert dict((n, c) for n, c, _ in api.trace_report())["get_contraction_ptrn_"] == 5
