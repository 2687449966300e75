"""ctypes wrapper of oracle/_ref/libaces4_ref.so: the REFERENCE's own C++ (compiled in place from /root/reference by
oracle/Makefile, target `ref`) behind the C ABI of oracle/ref_shim/aces4_ref_shim.cpp.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and by scripts/make_ref_golden.py, which turns its answers into the
committed fixtures under tests/golden/).  The product package never imports it.

`available()` is False where neither the prebuilt library nor the reference checkout exists; the tests that need the
live library skip there and the committed fixtures (generated from it) still pin the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libaces4_ref.so")
REFERENCE_ROOT = os.environ.get("ACES4_REFERENCE", "/root/reference")
_LIB = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
MAX_RANK = 6
GET, PUT, PUT_ACCUMULATE = 0, 1, 2
FILL, SCALE, SCALE_AND_COPY, COPY_DATA, INCREMENT, ACCUMULATE = range(6)
K_INT, K_DOUBLE, K_STRING, K_INT_ARRAY, K_DOUBLE_ARRAY, K_SIZE_T = range(6)


def build(force=False):
    """(Re)build oracle/_ref from the reference checkout when it is present; returns the path or None."""
    have_src = os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "sip"))
    if have_src and (force or not os.path.exists(_SO) or _stale()):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE_ROOT}"])
    return _SO if os.path.exists(_SO) else None


def _stale():
    mine = [os.path.join(_HERE, "ref_shim", f) for f in ("aces4_ref_shim.cpp", "config.h")]
    mine += [os.path.join(_HERE, "tensor_dil_oracle.c"), os.path.join(_HERE, "Makefile")]
    return os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in mine)


def available():
    try:
        return build() is not None
    except Exception:
        return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref is not built and the reference checkout is absent")
        _LIB = C.CDLL(path)
        _LIB.aces4ref_last_error.restype = C.c_char_p
        for f in ("aces4ref_shape_num_elems", "aces4ref_block_consistency", "aces4ref_block_number", "aces4ref_setup_dump"):
            getattr(_LIB, f).restype = C.c_longlong
        assert _LIB.aces4ref_max_rank() == MAX_RANK
    return _LIB


def _ia(seq, n=None):
    seq = [int(x) for x in seq]
    if n is not None:
        seq = seq + [0] * (n - len(seq))
    return (C.c_int * max(1, len(seq)))(*seq)


def _dp(a):
    return a.ctypes.data_as(c_dbl_p)


def _ok(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().aces4ref_last_error().decode()}")


def shape_num_elems(ext):
    return int(lib().aces4ref_shape_num_elems(len(ext), _ia(ext)))


def block_op(op, d, src=None, x=0.0):
    """Block::fill/scale/scale_and_copy/copy_data_/increment_elements/accumulate_data on a copy of d; returns the result."""
    out = np.array(d, dtype=np.float64, order="F", copy=True)
    s = None if src is None else np.asfortranarray(src, dtype=np.float64)
    _ok(lib().aces4ref_block_op(op, out.ndim, _ia(out.shape), _dp(out), _dp(s) if s is not None else None, C.c_double(x)),
        "block_op")
    return out


def transpose_copy(src, permute):
    """Block::transpose_copy with the 0-based permute vector of Interpreter::permute_rhs_to_lhs:
    permute[i] = position in the destination of the source's dimension i."""
    src = np.asfortranarray(src, dtype=np.float64)
    dext = [0] * src.ndim
    for i, p in enumerate(permute):
        dext[p] = src.shape[i]
    dst = np.zeros(dext, order="F")
    _ok(lib().aces4ref_block_transpose_copy(src.ndim, _ia(src.shape), _ia(dext), _ia(permute, MAX_RANK), _dp(src), _dp(dst)),
        "transpose_copy")
    return dst


def extract_slice(t, s_ext, offsets):
    t = np.asfortranarray(t, dtype=np.float64)
    s = np.zeros(s_ext, order="F")
    _ok(lib().aces4ref_block_extract_slice(t.ndim, _ia(t.shape), _dp(t), _ia(s_ext), _ia(offsets), _dp(s)), "extract_slice")
    return s


def insert_slice(t, s, offsets):
    out = np.array(t, dtype=np.float64, order="F", copy=True)
    s = np.asfortranarray(s, dtype=np.float64)
    _ok(lib().aces4ref_block_insert_slice(out.ndim, _ia(out.shape), _dp(out), _ia(s.shape), _ia(offsets), _dp(s)), "insert_slice")
    return out


def block_id_compare(array_a, idx_a, array_b, idx_b):
    r = lib().aces4ref_block_id_compare(array_a, _ia(idx_a, MAX_RANK), array_b, _ia(idx_b, MAX_RANK))
    if r == -2:
        raise RuntimeError(lib().aces4ref_last_error().decode())
    return r


def block_consistency(ops, workers, sections):
    """index of the first access DistributedBlockConsistency::update_and_check_consistency rejects, or -1"""
    r = int(lib().aces4ref_block_consistency(len(ops), _ia(ops), _ia(workers), _ia(sections)))
    if r == -2:
        raise RuntimeError(lib().aces4ref_last_error().decode())
    return r


def block_number(nseg, lower, idx):
    """(ArrayTableEntry::block_number(idx), num2id(that number)) for an array whose i-th index has nseg[i] segments from lower[i]"""
    back = (C.c_int * MAX_RANK)()
    n = int(lib().aces4ref_block_number(len(nseg), _ia(nseg), _ia(lower), _ia(idx), back))
    if n < 0:
        raise RuntimeError(lib().aces4ref_last_error().decode())
    return n, list(back[: len(nseg)])


def setup_dump(path):
    need = int(lib().aces4ref_setup_dump(path.encode(), None, 0))
    if need < 0:
        raise RuntimeError(lib().aces4ref_last_error().decode())
    buf = C.create_string_buffer(need + 1)
    lib().aces4ref_setup_dump(path.encode(), buf, need + 1)
    return buf.value.decode(errors="replace")


def setup_segments(path, index_type):
    out = (C.c_int * 256)()
    n = lib().aces4ref_setup_segments(path.encode(), int(index_type), out, 256)
    if n < 0:
        raise RuntimeError(lib().aces4ref_last_error().decode())
    return list(out[:n])


def setup_predefined_int(path, name):
    v = C.c_int()
    _ok(lib().aces4ref_setup_predefined_int(path.encode(), name.encode(), C.byref(v)), "predefined_int")
    return v.value


def setup_predefined_scalar(path, name):
    v = C.c_double()
    _ok(lib().aces4ref_setup_predefined_scalar(path.encode(), name.encode(), C.byref(v)), "predefined_scalar")
    return v.value


def _pools(records):
    kinds, ints, dbls, strs, lens = [], [], [], [], []
    for kind, val in records:
        kinds.append(kind)
        lens.append(len(val) if kind in (K_INT_ARRAY, K_DOUBLE_ARRAY) else 0)
        if kind in (K_INT, K_SIZE_T):
            ints.append(int(val))
        elif kind == K_DOUBLE:
            dbls.append(float(val))
        elif kind == K_STRING:
            strs.append(val.encode())
        elif kind == K_INT_ARRAY:
            ints.extend(int(v) for v in val)
        elif kind == K_DOUBLE_ARRAY:
            dbls.extend(float(v) for v in val)
    return kinds, ints, dbls, strs, lens


def stream_write(path, records):
    """records: [(kind, value)] written with the reference's setup::BinaryOutputFile (io_utils.cpp)"""
    kinds, ints, dbls, strs, lens = _pools(records)
    _ok(lib().aces4ref_stream_write(path.encode(), len(kinds), _ia(kinds), (C.c_longlong * max(1, len(ints)))(*ints),
                                    (C.c_double * max(1, len(dbls)))(*dbls), (C.c_char_p * max(1, len(strs)))(*strs),
                                    _ia(lens)), "stream_write")


def stream_read(path, kinds, max_items=1 << 16):
    """read records of the given kinds with the reference's setup::BinaryInputFile; returns [(kind, value)]"""
    ints = (C.c_longlong * max_items)()
    dbls = (C.c_double * max_items)()
    sbuf = C.create_string_buffer(1 << 16)
    lens = (C.c_int * max(1, len(kinds)))()
    _ok(lib().aces4ref_stream_read(path.encode(), len(kinds), _ia(kinds), ints, dbls, sbuf, C.c_longlong(len(sbuf)), lens),
        "stream_read")
    strs = sbuf.value.decode().split("\n")
    out, ii, di, si = [], 0, 0, 0
    for r, kind in enumerate(kinds):
        if kind in (K_INT, K_SIZE_T):
            out.append((kind, int(ints[ii])))
            ii += 1
        elif kind == K_DOUBLE:
            out.append((kind, float(dbls[di])))
            di += 1
        elif kind == K_STRING:
            out.append((kind, strs[si]))
            si += 1
        elif kind == K_INT_ARRAY:
            out.append((kind, [int(v) for v in ints[ii: ii + lens[r]]]))
            ii += lens[r]
        else:
            out.append((kind, [float(v) for v in dbls[di: di + lens[r]]]))
            di += lens[r]
    return out
