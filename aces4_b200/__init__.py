"""aces4_b200 -- B200-native block-tensor backend for the Aces4 SIP runtime (one hot path).

The product is the C-ABI shared library ``aces4_b200/lib/libsipgpu.so`` (include/sipgpu.h), built from the
hand-written sm_100a CUDA in ``aces4_b200/csrc``.  This package is only a thin ctypes view of that ABI for
tests, benchmarks and Python drivers; it contains no arithmetic and NO CPU fallback: if the library is not
built, or no B200 is present, calls fail loudly.
"""
from .api import (  # noqa: F401
    SipGpuError,
    DeviceBlock,
    DistArray,
    lib,
    lib_path,
    build,
    init,
    sync,
    kernel_launches,
)
from . import api  # noqa: F401

__all__ = ["SipGpuError", "DeviceBlock", "DistArray", "lib", "lib_path", "build", "init", "sync", "kernel_launches",
           "api"]
