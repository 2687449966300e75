#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.json from the REFERENCE's own C++ (oracle/_ref/libaces4_ref.so, built in place from
/root/reference by `make -C oracle ref`).  Run in the build container, where the reference checkout exists:

    python scripts/make_ref_golden.py

The fixture travels with the repository; the reference library does not have to.  Every value in it was computed by
reference code: DistributedBlockConsistency::update_and_check_consistency, ArrayTableEntry::block_number / num2id,
BlockId ordering, Block::transpose_copy / extract_slice / insert_slice / elementwise loops (block.cpp; see the note in
oracle/ref_shim/aces4_ref_shim.cpp about the three Fortran kernels underneath), setup::SetupReader on the shipped .dat
files and setup::BinaryOutputFile for the worker-checkpoint byte stream.
"""
import glob
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import ref_cases as rc  # noqa: E402
from oracle import ref  # noqa: E402

OPS = {"fill": ref.FILL, "scale": ref.SCALE, "scale_and_copy": ref.SCALE_AND_COPY, "copy_data": ref.COPY_DATA,
       "increment": ref.INCREMENT, "accumulate": ref.ACCUMULATE}


def checkpoint_records(scalars, arrays):
    """the calls WorkerPersistentArrayManager::checkpoint_persistent makes (worker_persistent_array_manager.cpp:158-210)"""
    recs = [(ref.K_INT, len(scalars))]
    for k in sorted(scalars):                      # std::map order
        recs += [(ref.K_STRING, k), (ref.K_DOUBLE, scalars[k])]
    recs.append((ref.K_INT, len(arrays)))
    for label, dims, data in sorted(arrays):
        dims6 = list(dims) + [1] * (6 - len(dims))
        recs += [(ref.K_STRING, label), (ref.K_INT, 6), (ref.K_INT_ARRAY, dims6), (ref.K_DOUBLE_ARRAY, data)]
    return recs


def main():
    assert ref.available(), "oracle/_ref cannot be built here (no reference checkout)"
    out = {"generator": "scripts/make_ref_golden.py", "source": "oracle/_ref/libaces4_ref.so (reference C++ compiled in place)"}

    v = rc.VERDICT_CHARS
    out["consistency_exhaustive"] = {
        str(n): "".join(v[1 + ref.block_consistency(o, w, s)] for o, w, s in rc.consistency_exhaustive(n))
        for n in range(1, rc.CONSISTENCY_MAX_LEN + 1)}
    out["consistency_sectioned"] = [ref.block_consistency(o, w, s) for o, w, s in rc.consistency_sectioned()]

    out["block_number"] = []
    for nseg, lower, idx in rc.block_number_cases():
        num, back = ref.block_number(nseg, lower, idx)
        assert back == idx, (nseg, lower, idx, num, back)      # Sip.check_block_number_calc: the round trip is the identity
        out["block_number"].append(num)

    out["block_id_compare"] = [ref.block_id_compare(*c) for c in rc.block_id_cases()]

    out["transpose"] = []
    for k, (ext, perm) in enumerate(rc.transpose_cases()):
        out["transpose"].append(rc.digest(ref.transpose_copy(rc.seeded(ext, k), perm)))

    out["slice"] = []
    for k, (t_ext, s_ext, off) in enumerate(rc.slice_cases()):
        t = rc.seeded(t_ext, 1000 + k)
        s = ref.extract_slice(t, s_ext, off)
        t2 = ref.insert_slice(rc.seeded(t_ext, 2000 + k), s, off)
        out["slice"].append([rc.digest(s), rc.digest(t2)])

    out["elementwise"] = []
    for k, (op, ext, x) in enumerate(rc.elementwise_cases()):
        d, s = rc.seeded(ext, 3000 + k), rc.seeded(ext, 4000 + k)
        needs_src = op in ("scale_and_copy", "copy_data", "accumulate")
        out["elementwise"].append(rc.digest(ref.block_op(OPS[op], d, s if needs_src else None, x)))

    dat = {}
    for path in sorted(glob.glob(os.path.join(ref.REFERENCE_ROOT, "test", "*.dat"))):
        name = os.path.basename(path)
        try:
            segs = {str(t): ref.setup_segments(path, t) for t in (1001, 1002, 1003, 1004)}
        except RuntimeError:
            continue                                   # not a setup file (other .dat payloads live in test/ too)
        ints = {}
        for key in ("baocc", "eaocc", "bavirt", "eavirt", "norb", "nalpha_occupied", "nalpha_virtual"):
            try:
                ints[key] = ref.setup_predefined_int(path, key)
            except RuntimeError:
                pass
        dat[name] = {"segments": {k: s for k, s in segs.items() if s}, "ints": ints}
    out["dat"] = dat

    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "ref.ckpt")
        ref.stream_write(p, checkpoint_records(rc.CHECKPOINT_SCALARS, rc.CHECKPOINT_ARRAYS))
        out["checkpoint_hex"] = open(p, "rb").read().hex()
        ref.stream_write(p, checkpoint_records(rc.CHECKPOINT_SCALARS, []))
        out["checkpoint_scalars_only_hex"] = open(p, "rb").read().hex()

    dst = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
        f.write("\n")
    print("wrote", dst, os.path.getsize(dst), "bytes;", len(dat), ".dat files;",
          sum(len(s) for s in out["consistency_exhaustive"].values()), "exhaustive consistency verdicts")


if __name__ == "__main__":
    main()
