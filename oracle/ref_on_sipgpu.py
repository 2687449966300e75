"""The reference's own `sip::Block` code (block.cpp, compiled in place) running on the PRODUCT: oracle/_ref/
libaces4_ref_on_sipgpu.so is the same set of reference objects as libaces4_ref.so, but block.cpp's three calls into
`libtensordil` -- tensor_block_copy__, tensor_block_slice__, tensor_block_insert__ -- are resolved by libsipgpu.so
(INTEGRATION.md level 0: link-time replacement, no source change).

TEST INFRASTRUCTURE ONLY.  `run(cases)` executes in a child process (so that the oracle-linked variant of the same
reference classes, which other tests load, never shares a process with this one):
    {"op": "transpose", "ext": [...], "permute": [...], "seed": s}      Block::transpose_copy
    {"op": "extract", "t_ext": [...], "s_ext": [...], "off": [...], "seed": s}   Block::extract_slice
    {"op": "insert", ...same...}                                         Block::insert_slice
    {"op": "gpu_block", "ext": [...], "fill": x, "scale": f}             level 1 (libaces4_ref_l1_on_sipgpu.so, reference
                                                                         built with HAVE_CUDA): Block::new_gpu_block ->
                                                                         gpu_fill -> gpu_scale -> gpu_copy_data, read back;
                                                                         result = stack of (fresh, filled*scale, copy)
returns the result arrays, or raises WorkerFailed with the reference's own failure text (sip::fail) -- which is what
happens on a machine without a GPU: the product reports SIPGPU_E_NODEVICE through `ierr` and block.cpp:252 turns it into
sip::fail, as for any other backend error.
"""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libaces4_ref_on_sipgpu.so")
_SO_L1 = os.path.join(_HERE, "_ref", "libaces4_ref_l1_on_sipgpu.so")   # + HAVE_CUDA: the device half of sip::Block (level 1)
_SO_L2 = os.path.join(_HERE, "_ref", "libaces4_ref_l2_on_sipgpu.so")   # + include/sial_ops_device_aces4.hpp behind the reference's SialOps signatures (level 2)
REFERENCE_ROOT = os.environ.get("ACES4_REFERENCE", "/root/reference")


class WorkerFailed(RuntimeError):
    pass


def build(force=False):
    have_src = os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "sip"))
    product = os.path.join(_HERE, "..", "aces4_b200", "lib", "libsipgpu.so")
    shim = os.path.join(_HERE, "ref_shim", "aces4_ref_shim.cpp")
    if have_src and os.path.exists(product) and (force or not os.path.exists(_SO)
                                                 or os.path.getmtime(_SO) < os.path.getmtime(shim) or not os.path.exists(_SO_L2)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref_on_sipgpu", "ref_l1", "ref_l2", f"REF={REFERENCE_ROOT}"])
    return _SO if os.path.exists(_SO) and os.path.exists(_SO_L1) else None


def available():
    try:
        return build() is not None
    except Exception:
        return os.path.exists(_SO)


def run_level2_selftest(nseg=3, seg=4, reps=5, timeout=120):
    """INTEGRATION level 2 in a child process: SialOpsDeviceAces4 (reference signatures: BlockId&, Block::BlockPtr, pc) driven
    with real sip::BlockId / sip::Block objects.  Returns (blocks checked, mismatching elements, collective_sum seen by the
    scalar sink, read-back after put_replace / increment / scale) or raises WorkerFailed with the library's message."""
    if build() is None or not os.path.exists(_SO_L2):
        raise WorkerFailed("oracle/_ref/libaces4_ref_l2_on_sipgpu.so is not built and the reference checkout is absent")
    code = ("import ctypes as C, sys\n"
            f"L = C.CDLL({_SO_L2!r})\n"
            "L.aces4ref_l2_last_error.restype = C.c_char_p\n"
            "out = (C.c_double * 4)()\n"
            f"rc = L.aces4ref_l2_selftest({int(nseg)}, {int(seg)}, {int(reps)}, out)\n"
            "print('L2', rc, *list(out)) if rc == 0 else print('L2ERR', L.aces4ref_l2_last_error().decode())\n")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("L2")]
    if p.returncode != 0 or not line or line[-1].startswith("L2ERR"):
        raise WorkerFailed(f"exit {p.returncode}: {(line[-1] if line else p.stderr or p.stdout)[-600:]}")
    return tuple(float(x) for x in line[-1].split()[2:])


def seeded(shape, seed):
    return np.asfortranarray(np.random.default_rng(seed).uniform(-1.0, 1.0, size=tuple(shape)))


def run(cases, timeout=120):
    if build() is None:
        raise WorkerFailed("oracle/_ref/libaces4_ref_on_sipgpu.so is not built and the reference checkout is absent")
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "cases.pkl"), os.path.join(tmp, "out.npz")
        with open(fin, "wb") as f:
            pickle.dump(cases, f)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", fin, fout], capture_output=True, text=True,
                           timeout=timeout)
        if p.returncode != 0 or not os.path.exists(fout):
            raise WorkerFailed(f"exit {p.returncode}: {(p.stderr or p.stdout)[-600:]}")
        z = np.load(fout)
        return [np.asfortranarray(z[f"y{k}"]) for k in range(len(cases))]


def _worker(fin, fout):
    sys.path.insert(0, os.path.dirname(_HERE))
    from oracle import ref

    cases = pickle.load(open(fin, "rb"))
    level1 = any(c["op"] == "gpu_block" for c in cases)
    so = _SO_L1 if level1 else _SO
    ref._SO = so                       # the same ctypes wrappers over the product-linked build
    ref.build = lambda force=False: so
    out = {}
    for k, c in enumerate(cases):
        try:
            if c["op"] == "gpu_block":
                import ctypes as C

                n = int(np.prod(c["ext"]))
                bufs = np.full((3, n), np.nan)
                dp = C.POINTER(C.c_double)
                rc = ref.lib().aces4ref_gpu_block_roundtrip(len(c["ext"]), ref._ia(c["ext"]), C.c_double(c["fill"]),
                                                            C.c_double(c["scale"]), bufs[0].ctypes.data_as(dp),
                                                            bufs[1].ctypes.data_as(dp), bufs[2].ctypes.data_as(dp))
                if rc != 0:
                    raise RuntimeError(f"aces4ref_gpu_block_roundtrip rc={rc} (2: _gpu_allocate gave no device block): "
                                       + ref.lib().aces4ref_last_error().decode())
                y = bufs
            elif c["op"] == "transpose":
                y = ref.transpose_copy(seeded(c["ext"], c["seed"]), c["permute"])
            elif c["op"] == "extract":
                y = ref.extract_slice(seeded(c["t_ext"], c["seed"]), c["s_ext"], c["off"])
            else:
                y = ref.insert_slice(seeded(c["t_ext"], c["seed"]), seeded(c["s_ext"], c["seed"] + 1), c["off"])
        except RuntimeError as e:
            print(f"case {k} ({c['op']}): {e}", file=sys.stderr)
            sys.exit(4)
        out[f"y{k}"] = y
    np.savez(fout, **out)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--worker":
        _worker(sys.argv[2], sys.argv[3])
    else:
        print(__doc__)
