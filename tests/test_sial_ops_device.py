"""include/sial_ops_device.hpp -- the C++ SialOpsDevice class (third SialOps implementation, boundary 4): it must
compile against the public header alone (CPU) and reproduce the reference's put/get closed forms on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_sial_ops_device.cpp")


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
