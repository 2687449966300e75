"""ctypes wrapper of the CPU oracle (oracle/tensor_dil_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (aces4_b200) never imports this module.

Arrays are numpy float64 in *Fortran order* (column-major, first index fastest), which is the
block layout of the reference (src/sip/dynamic_data/block.h:67-227).
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("tensor_dil_oracle.c", "super_instr_oracle.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_tensor_size_by_shape_.restype = C.c_longlong
        _LIB.oracle_tensor_block_norm2__.restype = C.c_double
        _LIB.oracle_block_number.restype = C.c_longlong
    return _LIB


def _ia(seq):
    return (C.c_int * max(1, len(seq)))(*[int(x) for x in seq])


def _dp(a):
    return a.ctypes.data_as(c_dbl_p)


def use_openblas(nthreads=None):
    """Plug scipy's bundled OpenBLAS dgemm into the oracle (F90:762 calls BLAS dgemm).
    Returns a description string of the BLAS build."""
    import scipy  # noqa: F401  (only to locate scipy.libs)

    pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so")
    path = sorted(glob.glob(pat))[0]
    blas = C.CDLL(path, mode=C.RTLD_GLOBAL)
    if nthreads is not None:
        blas.scipy_openblas_set_num_threads(int(nthreads))
    blas.scipy_openblas_get_config.restype = C.c_char_p
    fn = C.cast(blas.scipy_dgemm_, C.c_void_p)
    lib().oracle_set_dgemm(fn)
    use_openblas._keep = blas
    return blas.scipy_openblas_get_config().decode()


def use_naive_gemm():
    lib().oracle_set_dgemm(C.c_void_p(0))


def get_contraction_ptrn(dlab, llab, rlab):
    """F90:87-142.  Returns (pattern list, ierr)."""
    aces = list(dlab) + list(llab) + list(rlab)
    out = (C.c_int * max(1, len(llab) + len(rlab)))()
    ierr = C.c_int(0)
    lib().oracle_get_contraction_ptrn_(C.byref(C.c_int(len(dlab))), C.byref(C.c_int(len(llab))),
                                       C.byref(C.c_int(len(rlab))), _ia(aces), out, C.byref(ierr))
    return list(out)[: len(llab) + len(rlab)], ierr.value


def determine_index_permutations(ptrn, lext, rext, dext):
    lo2n = (C.c_int * 33)()
    ro2n = (C.c_int * 33)()
    do2n = (C.c_int * 33)()
    dims = (C.c_longlong * 3)()
    tr = (C.c_int * 3)()
    lib().oracle_determine_index_permutations(_ia(ptrn), len(lext), _ia(lext), len(rext), _ia(rext), len(dext),
                                              _ia(dext), lo2n, ro2n, do2n, dims, tr)
    return (list(lo2n)[1:len(lext) + 1], list(ro2n)[1:len(rext) + 1], list(do2n)[1:len(dext) + 1], list(dims),
            [bool(x) for x in tr])


def block_copy(a, transp):
    """tensor_block_copy__: transp = [sign, new position of old dim 1, ...] (1-based)."""
    a = np.asfortranarray(a, dtype=np.float64)
    rank = a.ndim
    new_ext = [0] * rank
    for i in range(rank):
        new_ext[transp[i + 1] - 1] = a.shape[i]
    out = np.empty(new_ext, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().oracle_tensor_block_copy__(C.byref(C.c_int(8)), C.byref(C.c_int(rank)), _ia(a.shape), _ia(transp), _dp(a),
                                     _dp(out), C.byref(ierr))
    assert ierr.value == 0, ierr.value
    return out


def _prep(x):
    """(contiguous column-major array, rank, shape); numpy promotes 0-d to 1-d in asfortranarray."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 0:
        return x.reshape(1).copy(), 0, ()
    return np.asfortranarray(x), x.ndim, x.shape


def block_contract(ptrn, L, R, dext):
    """tensor_block_contract__ (assign semantics).  Returns (D, ierr)."""
    L, lrank, lshape = _prep(L)
    R, rrank, rshape = _prep(R)
    D = np.full(tuple(dext) if len(dext) else (), np.nan, dtype=np.float64, order="F")
    if D.ndim == 0:
        D = np.full((1,), np.nan)
    ierr = C.c_int(0)
    lib().oracle_tensor_block_contract__(C.byref(C.c_int(8)), _ia(ptrn), _dp(L), C.byref(C.c_int(lrank)),
                                         _ia(lshape), _dp(R), C.byref(C.c_int(rrank)), _ia(rshape), _dp(D),
                                         C.byref(C.c_int(len(dext))), _ia(dext), C.byref(ierr))
    return D, ierr.value


def contract_labels(dlab, dext, llab, L, rlab, R):
    """D[dlab] = L[llab]*R[rlab] (interpreter.cpp:1210-1262).  Returns (D, ierr)."""
    L = np.asfortranarray(L, dtype=np.float64)
    R = np.asfortranarray(R, dtype=np.float64)
    D = np.full(tuple(dext) if len(dext) else (1,), np.nan, dtype=np.float64, order="F")
    ierr = lib().oracle_block_contract_labels(len(dlab), _ia(dext), _ia(dlab), _dp(D), L.ndim if len(llab) else 0,
                                              _ia(L.shape), _ia(llab), _dp(L), R.ndim if len(rlab) else 0,
                                              _ia(R.shape), _ia(rlab), _dp(R))
    return D, ierr


def permute_labels(lhs_labels, rhs_labels, rhs):
    """lhs[lhs_labels] = rhs[rhs_labels] (block_permute_op)."""
    rhs = np.asfortranarray(rhs, dtype=np.float64)
    ext = {lab: e for lab, e in zip(rhs_labels, rhs.shape)}
    out = np.empty([ext[lab] for lab in lhs_labels], dtype=np.float64, order="F")
    ierr = lib().oracle_block_permute(rhs.ndim, _ia(rhs.shape), _ia(lhs_labels), _ia(rhs_labels), _dp(rhs), _dp(out))
    assert ierr == 0, ierr
    return out


def block_add(t0, t1, fac):
    t0 = np.asfortranarray(t0, dtype=np.float64)
    ierr = C.c_int(0)
    lib().oracle_tensor_block_add__(C.byref(C.c_int(8)), C.byref(C.c_int(t0.ndim)), _ia(t0.shape), _dp(t0),
                                    _dp(np.asfortranarray(t1)), C.byref(C.c_double(fac)), C.byref(ierr))
    return t0, ierr.value


def block_slice(t, s_ext, beg):
    t = np.asfortranarray(t, dtype=np.float64)
    s = np.empty(s_ext, dtype=np.float64, order="F")
    ierr = C.c_int(0)
    lib().oracle_tensor_block_slice__(C.byref(C.c_int(8)), C.byref(C.c_int(t.ndim)), _dp(t), _ia(t.shape), _dp(s),
                                      _ia(s_ext), _ia(beg), C.byref(ierr))
    return s, ierr.value


def block_insert(t, s, beg):
    t = np.asfortranarray(t, dtype=np.float64)
    s = np.asfortranarray(s, dtype=np.float64)
    ierr = C.c_int(0)
    lib().oracle_tensor_block_insert__(C.byref(C.c_int(8)), C.byref(C.c_int(t.ndim)), _dp(t), _ia(t.shape), _dp(s),
                                       _ia(s.shape), _ia(beg), C.byref(ierr))
    return t, ierr.value


def block_norm2(t):
    t = np.asfortranarray(t, dtype=np.float64)
    ierr = C.c_int(0)
    return lib().oracle_tensor_block_norm2__(C.byref(C.c_int(8)), _dp(t), C.byref(C.c_int(t.ndim)), _ia(t.shape),
                                             C.byref(ierr))


def fill_cyclic(shape, start):
    a = np.empty(shape, dtype=np.float64, order="F")
    lib().oracle_fill_block_cyclic(_dp(a), C.c_longlong(a.size), C.c_double(start))
    return a


def fill_sequential(shape, start):
    a = np.empty(shape, dtype=np.float64, order="F")
    lib().oracle_fill_block_sequential(_dp(a), C.c_longlong(a.size), C.c_double(start))
    return a


def block_number(nseg, lower, idx):
    return lib().oracle_block_number(len(nseg), _ia(nseg), _ia(lower), _ia(idx))


def block_num2id(nseg, lower, num):
    out = (C.c_int * len(nseg))()
    lib().oracle_block_num2id(len(nseg), _ia(nseg), _ia(lower), C.c_longlong(num), out)
    return list(out)


def block_owner(num, nowners):
    return lib().oracle_block_owner(C.c_longlong(num), nowners)


def fill_hash(shape, seed, tag, scale=1.0):
    a = np.empty(shape, dtype=np.float64, order="F")
    lib().oracle_fill_hash(_dp(a), C.c_longlong(a.size), C.c_ulonglong(seed), C.c_ulonglong(tag), C.c_double(scale))
    return a


# ----------------------------------------------------------------------------------------------------
# elementwise CC super-instructions (oracle/super_instr_oracle.c); blocks are modified IN PLACE like the reference
# ----------------------------------------------------------------------------------------------------
def si_energy_denominator_rhf(block, index_values, fock, moa_seg_ranges):
    fock = np.asfortranarray(fock, dtype=np.float64)
    ierr = lib().oracle_si_energy_denominator_rhf(block.ndim, _ia(index_values), _ia(block.shape), _dp(block), fock.ndim,
                                                  fock.shape[0], _dp(fock), _ia(moa_seg_ranges))
    return ierr


def si_energy_ty_denominator_rhf(block, index_values, fock, shift, moa_seg_ranges):
    fock = np.asfortranarray(fock, dtype=np.float64)
    return lib().oracle_si_energy_ty_denominator_rhf(block.ndim, _ia(index_values), _ia(block.shape), _dp(block), fock.shape[0],
                                                     _dp(fock), C.c_double(shift), _ia(moa_seg_ranges))


def si_stripi(x, iv0, y_shape, iv1, moa_seg_ranges):
    y = np.zeros(y_shape, order="F")
    ierr = lib().oracle_si_stripi(x.ndim, _ia(iv0), _ia(x.shape), _dp(x), _ia(iv1), _ia(y_shape), _dp(y), _ia(moa_seg_ranges))
    return y, ierr


def si_anti_symm_o(block, index_values, moa_seg_ranges):
    return lib().oracle_si_anti_symm_o(block.ndim, _ia(index_values), _ia(block.shape), _dp(block), _ia(moa_seg_ranges))


def si_anti_symm_v(block, index_values, moa_seg_ranges):
    return lib().oracle_si_anti_symm_v(block.ndim, _ia(index_values), _ia(block.shape), _dp(block), _ia(moa_seg_ranges))


def si_return_sval(block):
    out = C.c_double(0)
    ierr = lib().oracle_si_return_sval(block.ndim, _ia(block.shape), _dp(block), C.byref(out))
    return out.value, ierr


def si_invert_diagonal(a1, a2):
    return lib().oracle_si_invert_diagonal(a1.ndim, a2.ndim, _ia(a1.shape), _dp(a1), _dp(a2))


def si_invert_diagonal_asym(a1, index_values, a2, moa_seg_ranges):
    return lib().oracle_si_invert_diagonal_asym(a1.ndim, a2.ndim, _ia(index_values), _ia(a1.shape), _dp(a1), _dp(a2), _ia(moa_seg_ranges))


def si_return_diagonal_elements(block, index_values, moa_seg_ranges):
    return lib().oracle_si_return_diagonal_elements(block.ndim, _ia(index_values), _ia(block.shape), _dp(block), _ia(moa_seg_ranges))


def block_consistency(ops, workers, sections):
    """distributed_block_consistency.cpp:25-175 for ONE block: index of the first illegal operation, or -1."""
    n = len(ops)
    lib().oracle_block_consistency.restype = C.c_longlong
    return int(lib().oracle_block_consistency(C.c_longlong(n), _ia(ops), _ia(workers), _ia(sections)))


def lazy_gpu_transition(bits, op):
    """block_manager.cpp:340-441: (failed, new bits, action)"""
    nb, act = C.c_int(0), C.c_int(0)
    rc = lib().oracle_lazy_gpu_transition(int(bits), int(op), C.byref(nb), C.byref(act))
    return rc, nb.value, act.value
