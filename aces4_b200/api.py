"""ctypes view of libsipgpu.so (include/sipgpu.h).  No arithmetic happens in this file.

Blocks are numpy float64 arrays in *Fortran order* on the host (column-major, first index fastest -- the block
layout of the reference, src/sip/dynamic_data/block.h:67-227) and ``DeviceBlock`` handles on the device.
Function names mirror the reference interface they bind: the ``tensor_block_*`` host-pointer ABI of
tensor_ops_c_prototypes.h:41-178, the ``_gpu_*`` device ABI of gpu_super_instructions.h:26-126 and the
SialOpsParallel get/put/put_accumulate of sial_ops_parallel.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_dbl_pp = C.POINTER(c_dbl_p)

# every symbol include/sipgpu.h declares (checked by tests/test_abi_symbols.py against the header itself)
BOUNDARY1 = ["tensor_size_by_shape_", "get_contraction_ptrn_", "tensor_block_init__", "tensor_block_scale__",
             "tensor_block_norm2__", "tensor_block_slice__", "tensor_block_insert__", "tensor_block_add__",
             "tensor_block_copy__", "tensor_block_contract__"]
BOUNDARY2 = ["_init_gpu", "_finalize_gpu", "_gpu_allocate", "_gpu_free", "_gpu_host_to_device", "_gpu_device_to_host",
             "_gpu_device_to_device", "_gpu_double_memset", "_gpu_selfmultiply", "_gpu_axpy", "_gpu_permute",
             "_gpu_contract"]


class SipGpuError(RuntimeError):
    pass


def lib_path():
    # SIPGPU_LIB: development override (kernel-tuning experiments load an alternative build of the SAME sources)
    return os.environ.get("SIPGPU_LIB") or os.path.join(_HERE, "lib", "libsipgpu.so")


def build(force=False, verbose=False):
    """Compile aces4_b200/csrc for sm_100a into aces4_b200/lib/libsipgpu.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        subprocess.check_call(cmd + ["clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return lib_path()


def lib():
    """Load the shared library.  There is no fallback: a missing build is an error."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise SipGpuError(f"{p} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or `make -C aces4_b200/csrc`); there is no CPU fallback")
        L = C.CDLL(p)
        L.tensor_size_by_shape_.restype = C.c_longlong
        L.tensor_block_norm2__.restype = C.c_double
        L._gpu_allocate.restype = c_dbl_p
        L.sipgpu_block_alloc.restype = c_dbl_p
        L.sipgpu_block_alloc.argtypes = [C.c_longlong, C.c_int]
        L.sipgpu_last_error.restype = C.c_char_p
        L.sipgpu_stream.restype = C.c_void_p
        L.sipgpu_kernel_launches.restype = C.c_longlong
        L.sipgpu_host_alloc.restype = C.c_void_p
        L.sipgpu_host_alloc.argtypes = [C.c_size_t]
        L.sipgpu_host_free.argtypes = [C.c_void_p]
        L.sipgpu_pool_reserve.argtypes = [C.c_size_t]
        L.sipgpu_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.sipgpu_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.sipgpu_block_free.argtypes = [C.c_void_p]
        for name in ("fill", "scale", "increment"):
            getattr(L, "sipgpu_block_" + name).argtypes = [C.c_void_p, C.c_longlong, C.c_double]
        L.sipgpu_block_scale_and_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double]
        L.sipgpu_block_axpy.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double]
        L.sipgpu_block_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.sipgpu_block_add_sub.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_double]
        L.sipgpu_block_fill_hash.argtypes = [C.c_void_p, C.c_longlong, C.c_ulonglong, C.c_ulonglong, C.c_double]
        L.sipgpu_block_dot_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        L.sipgpu_layout_block_number.argtypes = [C.c_int, c_int_p, c_int_p]
        L.sipgpu_layout_block_number.restype = C.c_longlong
        L.sipgpu_layout_block_owner.argtypes = [C.c_longlong, C.c_int]
        L.sipgpu_array_local_base.argtypes = [C.c_void_p]
        L.sipgpu_array_local_base.restype = C.c_void_p
        L.sipgpu_block_norm2.argtypes = [C.c_void_p, C.c_longlong, c_dbl_p]
        L.sipgpu_block_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, c_dbl_p]
        L.sipgpu_block_slice.argtypes = [C.c_int, C.c_void_p, c_int_p, C.c_void_p, c_int_p, c_int_p]
        L.sipgpu_block_insert.argtypes = [C.c_int, C.c_void_p, c_int_p, C.c_void_p, c_int_p, c_int_p]
        L.sipgpu_block_permute.argtypes = [C.c_int, c_int_p, c_int_p, C.c_void_p, C.c_void_p]
        L.sipgpu_permute_batched.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_double]
        L.sipgpu_block_permute_labels.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p, C.c_void_p, C.c_void_p]
        L.sipgpu_block_contract.argtypes = [c_int_p, C.c_void_p, C.c_int, c_int_p, C.c_void_p, C.c_int, c_int_p,
                                            C.c_void_p, C.c_int, c_int_p, C.c_double, C.c_double]
        L.sipgpu_block_contract_labels.argtypes = [C.c_int, c_int_p, c_int_p, C.c_void_p, C.c_int, c_int_p, c_int_p,
                                                   C.c_void_p, C.c_int, c_int_p, c_int_p, C.c_void_p, C.c_double,
                                                   C.c_double]
        L.sipgpu_contract_batched.argtypes = [C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.sipgpu_contract_chained.argtypes = [C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p,
                                              c_int_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.sipgpu_plan_contract_chained.argtypes = [C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p,
                                                   c_int_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.sipgpu_plan_launch.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.sipgpu_plan_destroy.argtypes = [C.c_void_p]
        L.sipgpu_dgemm_tn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_double, C.c_void_p, C.c_int]
        L.sipgpu_set_tuning.argtypes = [C.c_char_p, C.c_double]
        L.sipgpu_dmma_peak_probe.argtypes = [C.c_int, c_dbl_p]
        L.sipgpu_copy_bw_probe.argtypes = [C.c_size_t, C.c_int, c_dbl_p]
        L._gpu_free.argtypes = [C.c_void_p]
        L._gpu_host_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L._gpu_device_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L._gpu_device_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L._gpu_double_memset.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L._gpu_selfmultiply.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L._gpu_axpy.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int]
        L._gpu_permute.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, C.c_void_p, C.c_int, c_int_p, c_int_p]
        L._gpu_contract.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, C.c_void_p, C.c_int, c_int_p, c_int_p,
                                    C.c_void_p, C.c_int, c_int_p, c_int_p]
        L.sipgpu_array_create.argtypes = [C.c_int, c_int_p, c_int_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sipgpu_array_destroy.argtypes = [C.c_void_p]
        L.sipgpu_array_export.argtypes = [C.c_void_p, C.c_void_p]
        L.sipgpu_array_attach.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.sipgpu_array_block_number.argtypes = [C.c_void_p, c_int_p]
        L.sipgpu_array_block_number.restype = C.c_longlong
        L.sipgpu_array_block_owner.argtypes = [C.c_void_p, C.c_longlong]
        L.sipgpu_array_block_size.argtypes = [C.c_void_p, c_int_p]
        L.sipgpu_array_block_size.restype = C.c_longlong
        L.sipgpu_array_block_ptr.argtypes = [C.c_void_p, c_int_p]
        L.sipgpu_array_block_ptr.restype = C.c_void_p
        L.sipgpu_array_get.argtypes = [C.c_void_p, c_int_p, C.c_void_p]
        L.sipgpu_array_put.argtypes = [C.c_void_p, c_int_p, C.c_void_p]
        L.sipgpu_array_put_accumulate.argtypes = [C.c_void_p, c_int_p, C.c_void_p]
        L.sipgpu_array_get_many.argtypes = [C.c_void_p, C.c_int, c_int_p, C.POINTER(C.c_void_p)]
        L.sipgpu_array_put_accumulate_many.argtypes = [C.c_void_p, C.c_int, c_int_p, C.POINTER(C.c_void_p)]
        L.sipgpu_array_fill_local.argtypes = [C.c_void_p, C.c_double]
        for name in ("put_initialize", "put_increment", "put_scale"):
            getattr(L, "sipgpu_array_" + name).argtypes = [C.c_void_p, c_int_p, C.c_double]
        L.sipgpu_array_local_bytes.argtypes = [C.c_void_p]
        L.sipgpu_array_local_bytes.restype = C.c_size_t
        L.sipgpu_persist_scalar.argtypes = [C.c_char_p, C.c_double]
        L.sipgpu_restore_scalar.argtypes = [C.c_char_p, c_dbl_p]
        L.sipgpu_persist_contiguous.argtypes = [C.c_char_p, C.c_void_p, C.c_int, c_int_p]
        L.sipgpu_restore_contiguous.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), c_int_p, c_int_p]
        L.sipgpu_persist_array.argtypes = [C.c_char_p, C.c_void_p]
        L.sipgpu_restore_array.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.sipgpu_persist_count.argtypes = [c_int_p, c_int_p, c_int_p]
        L.sipgpu_persist_checkpoint.argtypes = [C.c_char_p]
        L.sipgpu_persist_init_from_checkpoint.argtypes = [C.c_char_p]
        L.sipgpu_array_save.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.sipgpu_array_load.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.sipgpu_array_track_accesses.argtypes = [C.c_void_p, C.c_int]
        L.sipgpu_array_section_accesses.argtypes = [C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong), c_int_p]
        L.sipgpu_array_section_accesses.restype = C.c_longlong
        L.sipgpu_array_section_reset.argtypes = [C.c_void_p]
        L.sipgpu_consistency_validate.argtypes = [C.c_longlong, C.POINTER(C.c_longlong), c_int_p, c_int_p,
                                                  C.POINTER(C.c_longlong)]
        L.sipgpu_mirror_transition.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p]
        L.sipgpu_mirror_create.argtypes = [C.c_void_p, C.c_longlong, C.POINTER(C.c_void_p)]
        L.sipgpu_mirror_destroy.argtypes = [C.c_void_p]
        L.sipgpu_mirror_status.argtypes = [C.c_void_p]
        L.sipgpu_mirror_access.argtypes = [C.c_void_p, C.c_int]
        L.sipgpu_mirror_access.restype = C.c_void_p
        L.sipgpu_mirror_host_ptr.argtypes = [C.c_void_p]
        L.sipgpu_mirror_host_ptr.restype = C.c_void_p
        L.sipgpu_mirror_device_ptr.argtypes = [C.c_void_p]
        L.sipgpu_mirror_device_ptr.restype = C.c_void_p
        L.sipgpu_wl_begin.argtypes = [C.c_int]
        L.sipgpu_wl_set_limits.argtypes = [C.c_longlong, C.c_longlong]
        L.sipgpu_wl_set_idle_flush.argtypes = [C.c_longlong]
        L.sipgpu_wl_stats.argtypes = [C.POINTER(C.c_longlong)]
        L.sipgpu_wl_last_plan.argtypes = [C.c_int, c_int_p, c_int_p]
        L.sipgpu_wl_replays.restype = C.c_longlong
        _LIB = L
    return _LIB


def _check(rc, what="sipgpu call"):
    if rc != 0:
        msg = lib().sipgpu_last_error().decode(errors="replace")
        raise SipGpuError(f"{what} failed with code {rc}: {msg}")


def _ia(seq):
    seq = [int(x) for x in seq]
    return (C.c_int * max(1, len(seq)))(*seq)


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


def init(device=-1):
    _check(lib().sipgpu_init(int(device)), "sipgpu_init")
    return lib().sipgpu_device()


def sync():
    _check(lib().sipgpu_sync(), "sipgpu_sync")


def kernel_launches():
    return int(lib().sipgpu_kernel_launches())


def stream_handle():
    return lib().sipgpu_stream()


# ----------------------------------------------------------------------------------------------------
# Boundary 1: host-pointer libtensordil ABI (numpy in, numpy out).  Same call shapes as oracle/oracle.py.
# ----------------------------------------------------------------------------------------------------
def _prep(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 0:
        return x.reshape(1).copy(), 0, ()
    return np.asfortranarray(x), x.ndim, x.shape


def get_contraction_ptrn(dlab, llab, rlab):
    aces = list(dlab) + list(llab) + list(rlab)
    out = (C.c_int * max(1, len(llab) + len(rlab)))()
    ierr = C.c_int(0)
    lib().get_contraction_ptrn_(C.byref(C.c_int(len(dlab))), C.byref(C.c_int(len(llab))), C.byref(C.c_int(len(rlab))),
                                _ia(aces), out, C.byref(ierr))
    return list(out)[: len(llab) + len(rlab)], ierr.value


def tensor_size_by_shape(dims):
    ierr = C.c_int(0)
    n = lib().tensor_size_by_shape_(C.byref(C.c_int(len(dims))), _ia(dims), C.byref(ierr))
    return n, ierr.value


def tensor_block_contract(ptrn, L, R, dext):
    """tensor_block_contract__ with host buffers (assign semantics).  Returns (D, ierr)."""
    L, lrank, lshape = _prep(L)
    R, rrank, rshape = _prep(R)
    D = np.full(tuple(dext) if len(dext) else (1,), np.nan, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().tensor_block_contract__(C.byref(C.c_int(8)), _ia(ptrn), _hp(L), C.byref(C.c_int(lrank)), _ia(lshape), _hp(R),
                                  C.byref(C.c_int(rrank)), _ia(rshape), _hp(D), C.byref(C.c_int(len(dext))), _ia(dext),
                                  C.byref(ierr))
    return D, ierr.value


def tensor_block_copy(a, transp):
    a = np.asfortranarray(a, dtype=np.float64)
    rank = a.ndim
    new_ext = [0] * rank
    for i in range(rank):
        new_ext[transp[i + 1] - 1] = a.shape[i]
    out = np.full(new_ext, np.nan, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().tensor_block_copy__(C.byref(C.c_int(8)), C.byref(C.c_int(rank)), _ia(a.shape), _ia(transp), _hp(a), _hp(out),
                              C.byref(ierr))
    return out, ierr.value


def tensor_block_add(t0, t1, fac):
    t0 = np.array(t0, dtype=np.float64, order="F")
    t1 = np.asfortranarray(t1, dtype=np.float64)
    ierr = C.c_int(0)
    lib().tensor_block_add__(C.byref(C.c_int(8)), C.byref(C.c_int(t0.ndim)), _ia(t0.shape), _hp(t0), _hp(t1),
                             C.byref(C.c_double(fac)), C.byref(ierr))
    return t0, ierr.value


def tensor_block_init(shape, val):
    t = np.full(shape, np.nan, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().tensor_block_init__(C.byref(C.c_int(8)), _hp(t), C.byref(C.c_int(t.ndim)), _ia(t.shape),
                              C.byref(C.c_double(val)), C.byref(ierr))
    return t, ierr.value


def tensor_block_scale(t, fac):
    t = np.array(t, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().tensor_block_scale__(C.byref(C.c_int(8)), _hp(t), C.byref(C.c_int(t.ndim)), _ia(t.shape),
                               C.byref(C.c_double(fac)), C.byref(ierr))
    return t, ierr.value


def tensor_block_norm2(t):
    t = np.asfortranarray(t, dtype=np.float64)
    ierr = C.c_int(0)
    v = lib().tensor_block_norm2__(C.byref(C.c_int(8)), _hp(t), C.byref(C.c_int(t.ndim)), _ia(t.shape), C.byref(ierr))
    return v, ierr.value


def tensor_block_slice(t, s_ext, beg):
    t = np.asfortranarray(t, dtype=np.float64)
    s = np.full(s_ext, np.nan, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().tensor_block_slice__(C.byref(C.c_int(8)), C.byref(C.c_int(t.ndim)), _hp(t), _ia(t.shape), _hp(s), _ia(s_ext),
                               _ia(beg), C.byref(ierr))
    return s, ierr.value


def tensor_block_insert(t, s, beg):
    t = np.array(t, dtype=np.float64, order="F")
    s = np.asfortranarray(s, dtype=np.float64)
    ierr = C.c_int(0)
    lib().tensor_block_insert__(C.byref(C.c_int(8)), C.byref(C.c_int(t.ndim)), _hp(t), _ia(t.shape), _hp(s),
                                _ia(s.shape), _ia(beg), C.byref(ierr))
    return t, ierr.value


# ----------------------------------------------------------------------------------------------------
# Device-resident blocks (boundaries 2 and 3)
# ----------------------------------------------------------------------------------------------------
class DeviceBlock:
    """A dense column-major FP64 block resident in the device pool (the device half of sip::Block,
    block.h:198-204)."""

    def __init__(self, shape, zero=False, ptr=None, owned=True):
        self.shape = tuple(int(s) for s in shape)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.owned = owned
        if ptr is None:
            p = lib().sipgpu_block_alloc(self.size, 1 if zero else 0)
            if not p:
                raise SipGpuError("sipgpu_block_alloc failed: " + lib().sipgpu_last_error().decode())
            self.ptr = C.cast(p, C.c_void_p).value
        else:
            self.ptr = int(ptr)

    @property
    def rank(self):
        return len(self.shape)

    @classmethod
    def from_numpy(cls, a):
        a = np.asarray(a, dtype=np.float64)
        shape = a.shape
        a = np.asfortranarray(a) if a.ndim else a.reshape(1).copy()
        b = cls(shape)
        _check(lib().sipgpu_h2d(b.ptr, _hp(a), b.size), "sipgpu_h2d")
        sync()  # `a` may be a temporary
        return b

    def to_numpy(self):
        out = np.empty(self.shape if self.shape else (1,), dtype=np.float64, order="F")
        _check(lib().sipgpu_d2h(_hp(out), self.ptr, self.size), "sipgpu_d2h")
        return out if self.shape else out.reshape(())

    def free(self):
        if self.owned and self.ptr:
            _check(lib().sipgpu_block_free(self.ptr), "sipgpu_block_free")
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # elementwise block ops (block.cpp:132-268)
    def fill(self, v):
        _check(lib().sipgpu_block_fill(self.ptr, self.size, float(v)))
        return self

    def scale(self, f):
        _check(lib().sipgpu_block_scale(self.ptr, self.size, float(f)))
        return self

    def fill_hash(self, seed, tag, scale=1.0):
        _check(lib().sipgpu_block_fill_hash(self.ptr, self.size, int(seed), int(tag), float(scale)))
        return self

    def increment(self, d):
        _check(lib().sipgpu_block_increment(self.ptr, self.size, float(d)))
        return self

    def accumulate(self, other):
        _check(lib().sipgpu_block_accumulate(self.ptr, other.ptr, self.size))
        return self

    def axpy(self, other, f):
        _check(lib().sipgpu_block_axpy(self.ptr, other.ptr, self.size, float(f)))
        return self

    def scale_and_copy(self, other, f):
        _check(lib().sipgpu_block_scale_and_copy(self.ptr, other.ptr, self.size, float(f)))
        return self

    def set_add_sub(self, l, r, sign):
        _check(lib().sipgpu_block_add_sub(self.ptr, l.ptr, r.ptr, self.size, float(sign)))
        return self

    def norm2(self):
        out = C.c_double(0)
        _check(lib().sipgpu_block_norm2(self.ptr, self.size, C.byref(out)))
        return out.value

    def dot(self, other):
        out = C.c_double(0)
        _check(lib().sipgpu_block_dot(self.ptr, other.ptr, self.size, C.byref(out)))
        return out.value


def permute(src, transp, out=None):
    """out[new position of idx] = src[idx]; transp = [sign, new position of old dim 1, ...] (1-based)."""
    new_ext = [0] * src.rank
    for i in range(src.rank):
        new_ext[transp[i + 1] - 1] = src.shape[i]
    if out is None:
        out = DeviceBlock(new_ext)
    _check(lib().sipgpu_block_permute(src.rank, _ia(src.shape), _ia(transp), src.ptr, out.ptr), "sipgpu_block_permute")
    return out


def permute_batched(srcs, transp, outs, alpha=1.0, beta=0.0):
    """outs[i] = alpha * permuted(srcs[i]) + beta * outs[i] for n blocks of identical shape in ONE launch."""
    shape = srcs[0].shape
    _check(lib().sipgpu_permute_batched(len(srcs), len(shape), _ia(shape), _ia(transp), _ptr_array([b.ptr for b in srcs]),
                                        _ptr_array([b.ptr for b in outs]), float(alpha), float(beta)),
           "sipgpu_permute_batched")
    return outs


class BatchedPermute:
    """A prepared batch (pointer arrays marshalled once) that can be re-launched cheaply: the ctypes conversion of a
    thousand block pointers costs more than the kernel that permutes them."""

    def __init__(self, srcs, transp, outs):
        self.n, self.shape = len(srcs), tuple(srcs[0].shape)
        self.ext, self.transp = _ia(self.shape), _ia(transp)
        self.ins, self.outs = _ptr_array([b.ptr for b in srcs]), _ptr_array([b.ptr for b in outs])

    def launch(self, alpha=1.0, beta=0.0):
        _check(lib().sipgpu_permute_batched(self.n, len(self.shape), self.ext, self.transp, self.ins, self.outs, float(alpha),
                                            float(beta)), "sipgpu_permute_batched")


def permute_labels(lhs_labels, rhs_labels, rhs, out=None):
    """lhs[lhs_labels] = rhs[rhs_labels]  (block_permute_op)."""
    ext = {lab: e for lab, e in zip(rhs_labels, rhs.shape)}
    if out is None:
        out = DeviceBlock([ext[lab] for lab in lhs_labels])
    _check(lib().sipgpu_block_permute_labels(rhs.rank, _ia(rhs.shape), _ia(lhs_labels), _ia(rhs_labels), rhs.ptr,
                                             out.ptr), "sipgpu_block_permute_labels")
    return out


def contract(ptrn, L, R, dext, out=None, alpha=1.0, beta=0.0):
    if out is None:
        out = DeviceBlock(dext)
    _check(lib().sipgpu_block_contract(_ia(ptrn), L.ptr, L.rank, _ia(L.shape), R.ptr, R.rank, _ia(R.shape), out.ptr,
                                       len(dext), _ia(dext), float(alpha), float(beta)), "sipgpu_block_contract")
    return out


def contract_sliced(ptrn, L, lext, lbeg, R, rext, rbeg, dext, out=None, dbeg=None, alpha=1.0, beta=0.0):
    """D = alpha * L[slice] * R[slice] + beta * D where L / R (/ out when dbeg is given) are PARENT arrays and
    (ext, beg) select the block inside them (beg None: the operand is a dense block of extents ext)."""
    if out is None:
        out = DeviceBlock(dext)
    def par(blk, beg):
        return (_ia(blk.shape), _ia(beg)) if beg is not None else (None, None)
    lp, lb = par(L, lbeg)
    rp, rb = par(R, rbeg)
    dp, db = par(out, dbeg)
    _check(lib().sipgpu_block_contract_sliced(_ia(ptrn), C.c_void_p(L.ptr), len(lext), _ia(lext), lp, lb, C.c_void_p(R.ptr),
                                              len(rext), _ia(rext), rp, rb, C.c_void_p(out.ptr), len(dext), _ia(dext), dp, db,
                                              C.c_double(alpha), C.c_double(beta)), "sipgpu_block_contract_sliced")
    return out


def contract_labels(dlab, dext, llab, L, rlab, R, out=None, alpha=1.0, beta=0.0):
    """D[dlab] = alpha * L[llab]*R[rlab] + beta*D  (handle_contraction; reference op is alpha=1, beta=0)."""
    if out is None:
        out = DeviceBlock(dext)
    _check(lib().sipgpu_block_contract_labels(len(dlab), _ia(dext), _ia(dlab), out.ptr, len(llab), _ia(L.shape),
                                              _ia(llab), L.ptr, len(rlab), _ia(R.shape), _ia(rlab), R.ptr,
                                              float(alpha), float(beta)), "sipgpu_block_contract_labels")
    return out


def _ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
    return arr


def contract_batched(ptrn, Ls, Rs, Ds, alpha=1.0, beta=0.0):
    """One launch (per kernel variant) for a work-list of blocks sharing a pattern; extents may differ."""
    n = len(Ls)
    lrank, rrank, drank = Ls[0].rank, Rs[0].rank, Ds[0].rank
    lext = np.array([b.shape for b in Ls], dtype=np.int32).reshape(n, lrank)
    rext = np.array([b.shape for b in Rs], dtype=np.int32).reshape(n, rrank)
    dext = np.array([b.shape for b in Ds], dtype=np.int32).reshape(n, drank)
    _check(lib().sipgpu_contract_batched(n, _ia(ptrn), lrank, rrank, drank, lext.ctypes.data_as(c_int_p),
                                         rext.ctypes.data_as(c_int_p), dext.ctypes.data_as(c_int_p),
                                         _ptr_array([b.ptr for b in Ls]), _ptr_array([b.ptr for b in Rs]),
                                         _ptr_array([b.ptr for b in Ds]), float(alpha), float(beta)),
           "sipgpu_contract_batched")


class BatchedContraction:
    """A prepared work-list (pointer and extent arrays marshalled once) that can be re-launched cheaply."""

    def __init__(self, ptrn, lshapes, rshapes, dshapes, lptrs, rptrs, dptrs, chain_start=None):
        """Without chain_start: n independent blocks.  With chain_start (n+1 offsets into lptrs/rptrs): destination i
        is the sum over the operand pairs chain_start[i]..chain_start[i+1]-1 (sipgpu_contract_chained)."""
        self.n = len(dptrs)
        self.chain = None if chain_start is None else np.ascontiguousarray(chain_start, dtype=np.int32)
        self.ptrn = _ia(ptrn)
        self.lrank, self.rrank, self.drank = len(lshapes[0]), len(rshapes[0]), len(dshapes[0])
        self.lext = np.ascontiguousarray(lshapes, dtype=np.int32)
        self.rext = np.ascontiguousarray(rshapes, dtype=np.int32)
        self.dext = np.ascontiguousarray(dshapes, dtype=np.int32)
        self.L, self.R, self.D = _ptr_array(lptrs), _ptr_array(rptrs), _ptr_array(dptrs)
        self.plan = None

    def launch(self, alpha=1.0, beta=0.0, prepared=True):
        """prepared (default): the work-list is marshalled once into a resident plan (sipgpu_plan_*) and every later launch
        replays it without host-side work; prepared=False goes through sipgpu_contract_chained / _batched every time."""
        if prepared:
            if self.plan is None:
                h = C.c_void_p(0)
                _check(lib().sipgpu_plan_contract_chained(self.n, self.ptrn, self.lrank, self.rrank, self.drank,
                                                          self.lext.ctypes.data_as(c_int_p), self.rext.ctypes.data_as(c_int_p),
                                                          self.dext.ctypes.data_as(c_int_p),
                                                          self.chain.ctypes.data_as(c_int_p) if self.chain is not None else None,
                                                          self.L, self.R, self.D, C.byref(h)), "sipgpu_plan_contract_chained")
                self.plan = h
            _check(lib().sipgpu_plan_launch(self.plan, float(alpha), float(beta)), "sipgpu_plan_launch")
            return
        if self.chain is not None:
            _check(lib().sipgpu_contract_chained(self.n, self.ptrn, self.lrank, self.rrank, self.drank,
                                                 self.lext.ctypes.data_as(c_int_p), self.rext.ctypes.data_as(c_int_p),
                                                 self.dext.ctypes.data_as(c_int_p), self.chain.ctypes.data_as(c_int_p),
                                                 self.L, self.R, self.D, float(alpha), float(beta)),
                   "sipgpu_contract_chained")
            return
        _check(lib().sipgpu_contract_batched(self.n, self.ptrn, self.lrank, self.rrank, self.drank,
                                             self.lext.ctypes.data_as(c_int_p), self.rext.ctypes.data_as(c_int_p),
                                             self.dext.ctypes.data_as(c_int_p), self.L, self.R, self.D, float(alpha),
                                             float(beta)), "sipgpu_contract_batched")

    def destroy(self):
        if self.plan is not None:
            lib().sipgpu_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------
# Boundary 3c: deferred op stream (worklist.cu) -- record the per-block calls of a pardo body, launch them batched
# ----------------------------------------------------------------------------------------------------
WL_STAT_NAMES = ("recorded", "scheduled", "levels", "launches", "fused_accumulates", "chains", "chain_pairs",
                 "temps_elided", "flushes")


def wl_begin(dry=False):
    _check(lib().sipgpu_wl_begin(1 if dry else 0), "sipgpu_wl_begin")


def wl_flush():
    _check(lib().sipgpu_wl_flush(), "sipgpu_wl_flush")


def wl_end():
    _check(lib().sipgpu_wl_end(), "sipgpu_wl_end")
    return wl_stats()


def wl_stats():
    out = (C.c_longlong * 9)()
    _check(lib().sipgpu_wl_stats(out), "sipgpu_wl_stats")
    return dict(zip(WL_STAT_NAMES, [int(x) for x in out]))


def wl_replays():
    """flushes of the open / last recording that replayed a captured stream instead of scheduling it"""
    return int(lib().sipgpu_wl_replays())


def wl_last_plan():
    """(level, unit) per recorded op of the most recent flush, in program order."""
    n = lib().sipgpu_wl_last_plan(0, None, None)
    lv, un = (C.c_int * max(1, n))(), (C.c_int * max(1, n))()
    lib().sipgpu_wl_last_plan(n, lv, un)
    return list(lv)[:n], list(un)[:n]


class recording:
    """`with api.recording(): ...` -- every asynchronous block op issued inside is deferred and batched at exit."""

    def __init__(self, dry=False):
        self.dry, self.stats = dry, None

    def __enter__(self):
        wl_begin(self.dry)
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.stats = wl_end()
        else:
            lib().sipgpu_wl_end()
        return False


def dgemm_tn(m, n, k, A, lda, B, ldb, Cblk, ldc, alpha=1.0, beta=0.0):
    _check(lib().sipgpu_dgemm_tn(m, n, k, float(alpha), A.ptr, lda, B.ptr, ldb, float(beta), Cblk.ptr, ldc),
           "sipgpu_dgemm_tn")


def trace(on=True, reset=True):
    """per-entry-point call counts / host time (sipgpu_trace_*), the counterpart of the reference's Tracer"""
    lib().sipgpu_trace_enable(1 if on else 0)
    if reset:
        lib().sipgpu_trace_reset()


def trace_report(cap=256):
    """[(entry point, calls, host seconds)], most host time first"""
    names, calls, secs = (C.c_char_p * cap)(), (C.c_longlong * cap)(), (C.c_double * cap)()
    n = min(cap, lib().sipgpu_trace_report(cap, names, calls, secs))
    return [(names[i].decode(), int(calls[i]), float(secs[i])) for i in range(n)]


def set_tuning(key, value):
    """launch-policy knob (sipgpu.h: sipgpu_set_tuning), e.g. set_tuning("lowint_max_intensity", -1) for the tile kernel only"""
    _check(lib().sipgpu_set_tuning(key.encode(), float(value)), "sipgpu_set_tuning")


def dmma_peak_probe(iters=20000):
    out = C.c_double(0)
    _check(lib().sipgpu_dmma_peak_probe(int(iters), C.byref(out)), "sipgpu_dmma_peak_probe")
    return out.value


def copy_bw_probe(nbytes=1 << 30, reps=10):
    out = C.c_double(0)
    _check(lib().sipgpu_copy_bw_probe(int(nbytes), int(reps), C.byref(out)), "sipgpu_copy_bw_probe")
    return out.value


def slice_block(t, s_ext, beg):
    s = DeviceBlock(s_ext)
    _check(lib().sipgpu_block_slice(t.rank, t.ptr, _ia(t.shape), s.ptr, _ia(s_ext), _ia(beg)), "sipgpu_block_slice")
    return s


def insert_block(t, s, beg):
    _check(lib().sipgpu_block_insert(t.rank, t.ptr, _ia(t.shape), s.ptr, _ia(s.shape), _ia(beg)), "sipgpu_block_insert")
    return t


# ----------------------------------------------------------------------------------------------------
# Boundary 3b: elementwise CC super-instructions on device blocks (reference calling convention)
# ----------------------------------------------------------------------------------------------------
def set_predefined_int_array(name, values):
    _check(lib().sipgpu_set_predefined_int_array(name.encode(), len(values), _ia(values)), "sipgpu_set_predefined_int_array")


def _si_arg(block, index_values=None):
    """The 6-tuple (array_slot, rank, index_values, size, extents, data) of one super-instruction argument."""
    rank = block.rank if block.shape != () else 0
    iv = list(index_values) if index_values is not None else [1] * max(rank, 1)
    return [C.byref(C.c_int(0)), C.byref(C.c_int(rank)), _ia(iv), C.byref(C.c_int(block.size)), _ia(block.shape or (1,)),
            C.c_void_p(block.ptr)]


def _si_call(fn, what, *blocks_and_ivs):
    args = []
    for blk, iv in blocks_and_ivs:
        args += _si_arg(blk, iv)
    ierr = C.c_int(0)
    rc = fn(*args, C.byref(ierr))
    assert rc == ierr.value
    return rc


def si_energy_denominator_rhf(block, index_values, fock):
    return _si_call(lib().sipgpu_si_energy_denominator_rhf, "energy_denominator_rhf", (block, index_values), (fock, None))


def si_energy_ty_denominator_rhf(block, index_values, fock, shift_block):
    return _si_call(lib().sipgpu_si_energy_ty_denominator_rhf, "energy_ty_denominator_rhf", (block, index_values), (fock, None),
                    (shift_block, None))


def si_stripi(x, iv0, y, iv1):
    return _si_call(lib().sipgpu_si_stripi, "stripi", (x, iv0), (y, iv1))


def si_anti_symm_o(block, index_values):
    return _si_call(lib().sipgpu_si_anti_symm_o, "anti_symm_o", (block, index_values))


def si_anti_symm_v(block, index_values):
    return _si_call(lib().sipgpu_si_anti_symm_v, "anti_symm_v", (block, index_values))


def si_return_sval(block, scalar_block):
    args = _si_arg(block) + [C.byref(C.c_int(0)), C.byref(C.c_int(0)), _ia([1]), C.byref(C.c_int(1)), _ia([1]),
                             C.c_void_p(scalar_block.ptr)]
    ierr = C.c_int(0)
    return lib().sipgpu_si_return_sval(*args, C.byref(ierr))


def si_invert_diagonal(a1, a2):
    return _si_call(lib().sipgpu_si_invert_diagonal, "invert_diagonal", (a1, None), (a2, None))


def si_invert_diagonal_asym(a1, index_values, a2):
    return _si_call(lib().sipgpu_si_invert_diagonal_asym, "invert_diagonal_asym", (a1, index_values), (a2, index_values))


def si_return_diagonal_elements(block, index_values):
    return _si_call(lib().sipgpu_si_return_diagonal_elements, "return_diagonal_elements", (block, index_values))


# ----------------------------------------------------------------------------------------------------
# Host/device coherence of one block (BlockManager::lazy_gpu_*, block_manager.cpp:340-441)
# ----------------------------------------------------------------------------------------------------
ON_HOST, ON_GPU, DIRTY_ON_HOST, DIRTY_ON_GPU = 1, 2, 4, 8
READ_ON_DEVICE, WRITE_ON_DEVICE, UPDATE_ON_DEVICE, READ_ON_HOST, WRITE_ON_HOST, UPDATE_ON_HOST = range(6)


def mirror_transition(bits, op):
    """(rc, new bits, action) of the pure lazy_gpu_* state transition; host-only"""
    nb, act = C.c_int(0), C.c_int(0)
    rc = lib().sipgpu_mirror_transition(int(bits), int(op), C.byref(nb), C.byref(act))
    return rc, nb.value, act.value


class MirroredBlock:
    """A block with a host copy (numpy, Fortran order) and/or a device copy kept coherent by the lazy_gpu_* rules."""

    def __init__(self, host_array=None, shape=None):
        self.host = None if host_array is None else np.asfortranarray(host_array, dtype=np.float64)
        self.shape = tuple(self.host.shape) if self.host is not None else tuple(shape)
        self.size = int(np.prod(self.shape))
        h = C.c_void_p(0)
        _check(lib().sipgpu_mirror_create(_hp(self.host) if self.host is not None else None, self.size, C.byref(h)),
               "sipgpu_mirror_create")
        self.h = h

    def status(self):
        return lib().sipgpu_mirror_status(self.h)

    def on_device(self, op):
        p = lib().sipgpu_mirror_access(self.h, op)
        if not p:
            raise SipGpuError("mirror access failed: " + lib().sipgpu_last_error().decode(errors="replace"))
        return DeviceBlock(self.shape, ptr=p, owned=False)

    def on_host(self, op):
        p = lib().sipgpu_mirror_access(self.h, op)
        if not p:
            raise SipGpuError("mirror access failed: " + lib().sipgpu_last_error().decode(errors="replace"))
        buf = (C.c_double * self.size).from_address(p)
        return np.frombuffer(buf, dtype=np.float64).reshape(self.shape, order="F")

    def destroy(self):
        if self.h:
            lib().sipgpu_mirror_destroy(self.h)
            self.h = None


# ----------------------------------------------------------------------------------------------------
# Persistence (worker_persistent_array_manager.cpp:34-260): label registry + checkpoint file in the reference format
# ----------------------------------------------------------------------------------------------------
def persist_scalar(label, value):
    _check(lib().sipgpu_persist_scalar(label.encode(), float(value)), "sipgpu_persist_scalar")


def restore_scalar(label):
    v = C.c_double(0)
    _check(lib().sipgpu_restore_scalar(label.encode(), C.byref(v)), "sipgpu_restore_scalar")
    return v.value


def persist_contiguous(label, block):
    """Ownership of the pool block moves to the registry (the DeviceBlock handle is disowned)."""
    _check(lib().sipgpu_persist_contiguous(label.encode(), block.ptr, block.rank, _ia(block.shape)), "sipgpu_persist_contiguous")
    block.owned = False


def restore_contiguous(label):
    p, rank, ext = C.c_void_p(0), C.c_int(0), (C.c_int * 6)()
    _check(lib().sipgpu_restore_contiguous(label.encode(), C.byref(p), C.byref(rank), ext), "sipgpu_restore_contiguous")
    return DeviceBlock(tuple(ext)[: rank.value], ptr=p.value, owned=True)


def persist_counts():
    a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
    lib().sipgpu_persist_count(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def persist_checkpoint(path):
    _check(lib().sipgpu_persist_checkpoint(str(path).encode()), "sipgpu_persist_checkpoint")


def persist_init_from_checkpoint(path):
    _check(lib().sipgpu_persist_init_from_checkpoint(str(path).encode()), "sipgpu_persist_init_from_checkpoint")


# ----------------------------------------------------------------------------------------------------
# Boundary 4: distributed arrays
# ----------------------------------------------------------------------------------------------------
ACCESS_GET, ACCESS_PUT, ACCESS_PUT_ACCUMULATE = 1, 2, 4


def consistency_validate(entries):
    """entries: [(block number, access bits, worker)] of one barrier section, all ranks.  Raises SipGpuError(E_STATE)
    on an inconsistent block (distributed_block_consistency.cpp); host-only."""
    n = len(entries)
    b = (C.c_longlong * max(1, n))(*[int(e[0]) for e in entries])
    f = (C.c_int * max(1, n))(*[int(e[1]) for e in entries])
    w = (C.c_int * max(1, n))(*[int(e[2]) for e in entries])
    bad = C.c_longlong(-1)
    rc = lib().sipgpu_consistency_validate(n, b, f, w, C.byref(bad))
    if rc != 0:
        raise SipGpuError(f"inconsistent block {bad.value}: " + lib().sipgpu_last_error().decode(errors="replace"))


def layout_block_number(nseg, idx):
    """array_table.cpp:75-81 (LAST index fastest, segments 1-based); host-only."""
    return int(lib().sipgpu_layout_block_number(len(nseg), _ia(nseg), _ia(idx)))


def layout_block_owner(number, world):
    """data_distribution.cpp:74-82; host-only."""
    return int(lib().sipgpu_layout_block_owner(int(number), int(world)))


def partition_blocks(nseg, world):
    """owner -> list of 1-based block index tuples, in block-number order (the destinations a rank computes)."""
    out = [[] for _ in range(world)]
    for idx in np.ndindex(*nseg):
        idx1 = tuple(int(x) + 1 for x in idx)
        out[layout_block_owner(layout_block_number(nseg, idx1), world)].append(idx1)
    return out


class DistArray:
    """A distributed/served SIAL array: blocks owned block-cyclically by the ranks (one rank per GPU), slabs
    mapped across processes with CUDA IPC.  `exchange` is a callable all_gather(bytes) -> list[bytes]
    (torch.distributed.all_gather_object in the harness); with world == 1 it is not needed."""

    def __init__(self, seg_ext_per_index, my_rank=0, world=1, exchange=None, devices=None):
        self.seg_ext = [list(map(int, s)) for s in seg_ext_per_index]
        self.rank = len(self.seg_ext)
        self.my_rank, self.world = my_rank, world
        nseg = [len(s) for s in self.seg_ext]
        flat = [e for s in self.seg_ext for e in s]
        h = C.c_void_p(0)
        _check(lib().sipgpu_array_create(self.rank, _ia(nseg), _ia(flat), my_rank, world, C.byref(h)),
               "sipgpu_array_create")
        self.h = h
        if world > 1:
            buf = C.create_string_buffer(64)
            _check(lib().sipgpu_array_export(self.h, buf), "sipgpu_array_export")
            handles = exchange(buf.raw)
            for r, hb in enumerate(handles):
                if r != my_rank:
                    dev = devices[r] if devices else r
                    _check(lib().sipgpu_array_attach(self.h, r, C.create_string_buffer(hb, 64), dev),
                           "sipgpu_array_attach")

    def block_number(self, idx):
        return int(lib().sipgpu_array_block_number(self.h, _ia(idx)))

    def owner(self, idx):
        return int(lib().sipgpu_array_block_owner(self.h, self.block_number(idx)))

    def block_shape(self, idx):
        return tuple(self.seg_ext[i][idx[i] - 1] for i in range(self.rank))

    def block_ptr(self, idx):
        p = lib().sipgpu_array_block_ptr(self.h, _ia(idx))
        if not p:
            raise SipGpuError("block not mapped: " + lib().sipgpu_last_error().decode())
        return int(p)

    def block_view(self, idx):
        """DeviceBlock aliasing the owner's storage (peer memory when remote)."""
        return DeviceBlock(self.block_shape(idx), ptr=self.block_ptr(idx), owned=False)

    def get(self, idx, out=None):
        if out is None:
            out = DeviceBlock(self.block_shape(idx))
        _check(lib().sipgpu_array_get(self.h, _ia(idx), out.ptr), "sipgpu_array_get")
        return out

    def put(self, idx, blk):
        _check(lib().sipgpu_array_put(self.h, _ia(idx), blk.ptr), "sipgpu_array_put")

    def put_accumulate(self, idx, blk):
        _check(lib().sipgpu_array_put_accumulate(self.h, _ia(idx), blk.ptr), "sipgpu_array_put_accumulate")

    def section(self, idxs, blocks):
        """marshalled arguments of get_many / put_accumulate_many (build once, reuse per section)"""
        flat = [int(v) for ix in idxs for v in ix]
        return len(idxs), (C.c_int * len(flat))(*flat), (C.c_void_p * len(blocks))(*[b.ptr for b in blocks])

    def get_many(self, idxs=None, blocks=None, section=None):
        """one barrier section's worth of gets as one launch"""
        n, ia, pa = section or self.section(idxs, blocks)
        _check(lib().sipgpu_array_get_many(self.h, n, ia, pa), "sipgpu_array_get_many")

    def put_accumulate_many(self, idxs=None, blocks=None, section=None):
        n, ia, pa = section or self.section(idxs, blocks)
        _check(lib().sipgpu_array_put_accumulate_many(self.h, n, ia, pa), "sipgpu_array_put_accumulate_many")

    def put_initialize(self, idx, v):
        _check(lib().sipgpu_array_put_initialize(self.h, _ia(idx), float(v)), "sipgpu_array_put_initialize")

    def put_increment(self, idx, d):
        _check(lib().sipgpu_array_put_increment(self.h, _ia(idx), float(d)), "sipgpu_array_put_increment")

    def put_scale(self, idx, f):
        _check(lib().sipgpu_array_put_scale(self.h, _ia(idx), float(f)), "sipgpu_array_put_scale")

    def fill_local(self, v):
        _check(lib().sipgpu_array_fill_local(self.h, float(v)), "sipgpu_array_fill_local")

    def local_base(self):
        return int(lib().sipgpu_array_local_base(self.h) or 0)

    def local_bytes(self):
        return int(lib().sipgpu_array_local_bytes(self.h))

    def track_accesses(self, on=True):
        _check(lib().sipgpu_array_track_accesses(self.h, 1 if on else 0), "sipgpu_array_track_accesses")

    def section_accesses(self):
        """[(block number, access bits)] of what this rank did since the last barrier"""
        n = lib().sipgpu_array_section_accesses(self.h, 0, None, None)
        b, f = (C.c_longlong * max(1, n))(), (C.c_int * max(1, n))()
        lib().sipgpu_array_section_accesses(self.h, n, b, f)
        return list(zip(list(b)[:n], list(f)[:n]))

    def section_reset(self):
        _check(lib().sipgpu_array_section_reset(self.h), "sipgpu_array_section_reset")

    def save(self, data_path, index_path):
        _check(lib().sipgpu_array_save(self.h, str(data_path).encode(), str(index_path).encode()), "sipgpu_array_save")

    def load(self, data_path, index_path):
        _check(lib().sipgpu_array_load(self.h, str(data_path).encode(), str(index_path).encode()), "sipgpu_array_load")

    def persist(self, label):
        """set_persistent: the array object (and its HBM slab) moves to the label registry."""
        _check(lib().sipgpu_persist_array(label.encode(), self.h), "sipgpu_persist_array")
        self.h = None

    def restore(self, label):
        """restore_persistent into this (freshly declared, same-layout) array: adopts the persisted slab."""
        h = C.c_void_p(0)
        _check(lib().sipgpu_restore_array(label.encode(), C.byref(h)), "sipgpu_restore_array")
        if self.h:
            lib().sipgpu_array_destroy(self.h)
        self.h = h

    def destroy(self):
        if self.h:
            lib().sipgpu_array_destroy(self.h)
            self.h = None
