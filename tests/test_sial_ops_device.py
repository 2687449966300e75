"""include/sial_ops_device.hpp -- the C++ SialOpsDevice class (third SialOps implementation, boundary 4): it must
compile against the public header alone (CPU) and reproduce the reference's put/get closed forms on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_sial_ops_device.cpp")


# the public block-traffic methods of SialOpsParallel (src/sip/worker/sial_ops_parallel.h:47-73)
SIAL_OPS_METHODS = ["sip_barrier", "create_distributed", "restore_distributed", "delete_distributed", "get", "put_replace",
                    "put_accumulate", "put_initialize", "put_increment", "put_scale", "destroy_served", "request", "prequest",
                    "prepare", "prepare_accumulate", "collective_sum", "assert_same", "broadcast_static", "set_persistent",
                    "restore_persistent", "end_program"]


def test_method_set_covers_sial_ops_parallel():
    """every public method of the reference's SialOpsParallel has a same-named method in SialOpsDevice; where the
    reference checkout is present the list above is re-derived from its header"""
    import re

    mine = open(os.path.join(ROOT, "include", "sial_ops_device.hpp")).read()
    for name in SIAL_OPS_METHODS:
        assert re.search(r"\b" + name + r"\s*\(", mine), name
    ref = "/root/reference/src/sip/worker/sial_ops_parallel.h"
    if os.path.exists(ref):
        text = open(ref).read()
        public = text[text.index("void sip_barrier"): text.index("void end_program")] + "void end_program();"
        names = re.findall(r"^\s*(?:void|bool)\s+(\w+)\s*\(", public, flags=re.M)
        assert names == SIAL_OPS_METHODS, names


def test_header_compiles_standalone():
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), SRC])


@pytest.mark.gpu
def test_sial_ops_device_closed_forms(tmp_path):
    import aces4_b200 as sip

    lib_dir = os.path.dirname(sip.lib_path())
    exe = str(tmp_path / "test_sial_ops_device")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), SRC, "-L", lib_dir, "-lsipgpu",
                           f"-Wl,-rpath,{lib_dir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr
