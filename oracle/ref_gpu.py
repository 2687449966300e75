"""The REFERENCE's legacy CUDA backend (src/sip/cuda/gpu_super_instructions.cu, compiled unmodified for sm_100a into
oracle/_ref/libaces4_ref_gpu.so by `make -C oracle ref_gpu`) as a checker and a second baseline.

TEST INFRASTRUCTURE ONLY -- used by tests/test_gpu_vs_ref_cuda.py and scripts/ref_gpu_baseline.py; the product never
imports it.  The reference code calls exit(EXIT_FAILURE) on any CUDA error (cuda_check.h:37-43) and prints to stdout, so
it always runs in a CHILD process: `run_cases` ships a list of seeded cases to `python oracle/ref_gpu.py --worker`, which
generates the inputs (numpy default_rng(seed), column-major), runs `_gpu_contract` / `_gpu_permute` of the reference on
device buffers it allocated with the reference's `_gpu_allocate`, and returns the results (and wall time per call --
every reference call ends with cudaDeviceSynchronize, so host timing is device timing).

A case is a dict: {"kind": "contract", "y": (dims, labels), "x1": (dims, labels), "x2": (dims, labels), "seed": s}
               or {"kind": "permute",  "y": (dims, labels), "x1": (dims, labels), "seed": s}; optional "reps": n (timing).
Labels are arbitrary distinct ints (index-table slots in the reference, interpreter.cpp:2699-2798).
Limits of the reference code: blocks of at most 40 MB (fixed scratch buffers), rank <= 6.
"""
import ctypes as C
import os
import pickle
import subprocess
import sys
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libaces4_ref_gpu.so")
REFERENCE_ROOT = os.environ.get("ACES4_REFERENCE", "/root/reference")
MAX_BLOCK_DOUBLES = 40 * 1024 * 1024 // 8


class WorkerFailed(RuntimeError):
    pass


def build(force=False):
    src = os.path.join(REFERENCE_ROOT, "src", "sip", "cuda", "gpu_super_instructions.cu")
    shim = os.path.join(_HERE, "ref_shim", "aces4_ref_gpu_shim.cpp")
    if os.path.exists(src) and (force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(shim)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref_gpu", f"REF={REFERENCE_ROOT}"])
    return _SO if os.path.exists(_SO) else None


def available():
    try:
        return build() is not None
    except Exception:
        return os.path.exists(_SO)


def case_inputs(case):
    """the operands of a case, column-major, from its seed (same generator in the worker and in the test)"""
    rng = np.random.default_rng(case["seed"])
    x1 = np.asfortranarray(rng.uniform(-1.0, 1.0, size=tuple(case["x1"][0])))
    x2 = np.asfortranarray(rng.uniform(-1.0, 1.0, size=tuple(case["x2"][0]))) if case["kind"] == "contract" else None
    return x1, x2


def run_cases(cases, timeout=300):
    """-> (list of result arrays (column-major, shape = y dims), list of seconds per call)"""
    if build() is None:
        raise WorkerFailed("oracle/_ref/libaces4_ref_gpu.so is not built and the reference checkout is absent")
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "cases.pkl"), os.path.join(tmp, "out.npz")
        with open(fin, "wb") as f:
            pickle.dump(cases, f)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", fin, fout], capture_output=True,
                               text=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            raise WorkerFailed(f"reference CUDA worker timed out after {timeout} s") from None
        if p.returncode != 0 or not os.path.exists(fout):
            raise WorkerFailed(f"reference CUDA worker exited with {p.returncode}: {(p.stderr or p.stdout)[-800:]}")
        z = np.load(fout)
        return [np.asfortranarray(z[f"y{k}"]) for k in range(len(cases))], [float(t) for t in z["seconds"]]


# ---------------------------------------------------------------------------------------------------------------------
def _worker(fin, fout):
    cases = pickle.load(open(fin, "rb"))
    L = C.CDLL(_SO)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.aces4ref_gpu_allocate.restype = dp
    L.aces4ref_gpu_allocate.argtypes = [C.c_int]
    L.aces4ref_gpu_free.argtypes = [dp]
    L.aces4ref_gpu_host_to_device.argtypes = [dp, dp, C.c_int]
    L.aces4ref_gpu_device_to_host.argtypes = [dp, dp, C.c_int]
    L.aces4ref_gpu_permute.argtypes = [dp, C.c_int, ip, ip, dp, C.c_int, ip, ip]
    L.aces4ref_gpu_contract.argtypes = [dp, C.c_int, ip, ip, dp, C.c_int, ip, ip, dp, C.c_int, ip, ip]
    dev = L.aces4ref_gpu_init(int(os.environ.get("LOCAL_RANK", "0")))
    if dev < 0:
        print("reference _init_gpu found no device", file=sys.stderr)
        sys.exit(3)

    def ia(v):
        v = [int(x) for x in v] + [1] * (6 - len(v))      # the reference passes MAX_RANK-sized arrays
        return (C.c_int * len(v))(*v)

    def up(a):
        g = L.aces4ref_gpu_allocate(int(a.size))
        L.aces4ref_gpu_host_to_device(a.ctypes.data_as(dp), g, int(a.size))
        return g

    out, secs = {}, []
    for k, c in enumerate(cases):
        x1, x2 = case_inputs(c)
        ydims, yinds = c["y"]
        for a in (x1, x2):
            assert a is None or a.size <= MAX_BLOCK_DOUBLES, "block larger than the reference's 40 MB scratch buffers"
        y = np.zeros(tuple(ydims), order="F")
        assert y.size <= MAX_BLOCK_DOUBLES
        g1, gy = up(x1), L.aces4ref_gpu_allocate(int(y.size))
        g2 = up(x2) if x2 is not None else None
        reps = int(c.get("reps", 1))

        def call():
            if c["kind"] == "contract":
                L.aces4ref_gpu_contract(gy, len(ydims), ia(ydims), ia(yinds), g1, x1.ndim, ia(c["x1"][0]), ia(c["x1"][1]),
                                        g2, x2.ndim, ia(c["x2"][0]), ia(c["x2"][1]))
            else:
                L.aces4ref_gpu_permute(gy, len(ydims), ia(ydims), ia(yinds), g1, x1.ndim, ia(c["x1"][0]), ia(c["x1"][1]))

        if reps > 1:
            call()          # timing run: one untimed call first (cuBLAS workspace, module load)
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        secs.append((time.perf_counter() - t0) / reps)
        L.aces4ref_gpu_device_to_host(y.ctypes.data_as(dp), gy, int(y.size))
        out[f"y{k}"] = y
        for g in (g1, g2, gy):
            if g is not None:
                L.aces4ref_gpu_free(g)
    np.savez(fout, seconds=np.array(secs), **out)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--worker":
        _worker(sys.argv[2], sys.argv[3])
    else:
        print(__doc__)
