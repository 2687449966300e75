"""The oracle and the product's host logic against the REFERENCE's own C++.

Two layers, same cases (tests/ref_cases.py):
 * tests/golden/ref_vectors.json -- answers computed by oracle/_ref (the reference's block.cpp, block_id.cpp,
   distributed_block_consistency.cpp, array_table.cpp, index_table.cpp, setup_reader.cpp, io_utils.cpp compiled in place
   from /root/reference by `make -C oracle ref`; generator: scripts/make_ref_golden.py).  Always runs.
 * the live library, where it is present (this container, and any box the prebuilt oracle/_ref travelled to): the same
   comparisons plus cross-reads of files (product checkpoint read by the reference's BinaryInputFile, and the reverse).

What this pins on reference CODE rather than on a reading of it: the consistency state table, block numbering and its
inverse, BlockId order, the permute-vector convention of Block::transpose_copy (0-based "destination position of source
dimension i"), slice offsets, the elementwise block loops, the .dat segment tables and the checkpoint byte stream.
The Fortran loop nests themselves (tensor_dil_omp.F90) cannot be compiled here; they stay pinned by the reference's
known-answer tests restated in tests/test_oracle_golden.py.
"""
import json
import os
import struct

import numpy as np
import pytest

import ref_cases as rc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))


def _ref_or_none():
    try:
        from oracle import ref

        return ref if ref.available() else None
    except Exception:
        return None


REF = _ref_or_none()
live = pytest.mark.skipif(REF is None, reason="oracle/_ref not built and no reference checkout on this machine")


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


BITS = {0: 1, 1: 2, 2: 4}   # GET, PUT, PUT_ACCUMULATE -> SIPGPU_ACCESS_*


def library_accepts(sip, ops, workers, sections):
    """the product validates one barrier section at a time (sipgpu_consistency_validate)"""
    for sec in sorted(set(sections)):
        entries = [(7, BITS[o], w) for o, w, s in zip(ops, workers, sections) if s == sec]
        try:
            sip.consistency_validate(entries)
        except sip.SipGpuError:
            return False
    return True


# ---------------------------------------------------------------------------------------------------------------
# fixture layer
# ---------------------------------------------------------------------------------------------------------------
def test_consistency_table_against_reference_verdicts(oracle, sip):
    """all 66 429 sequences of <= 5 accesses by 3 workers: oracle == reference verdict (index of the first rejected access);
    product == reference on accept / reject (its rule is order-independent, so only the verdict is comparable)"""
    for n in range(1, rc.CONSISTENCY_MAX_LEN + 1):
        want = GOLD["consistency_exhaustive"][str(n)]
        cases = list(rc.consistency_exhaustive(n))
        assert len(want) == len(cases)
        for k, (ops, workers, sections) in enumerate(cases):
            ref_verdict = rc.VERDICT_CHARS.index(want[k]) - 1
            assert oracle.block_consistency(ops, workers, sections) == ref_verdict, (ops, workers)
            if n <= 4 or k % 7 == 0:
                assert library_accepts(sip, ops, workers, sections) == (ref_verdict == -1), (ops, workers)


def test_consistency_with_barrier_sections_against_reference_verdicts(oracle, sip):
    cases = list(rc.consistency_sectioned())
    assert len(cases) == len(GOLD["consistency_sectioned"])
    for (ops, workers, sections), want in zip(cases, GOLD["consistency_sectioned"]):
        assert oracle.block_consistency(ops, workers, sections) == want, (ops, workers, sections)
        assert library_accepts(sip, ops, workers, sections) == (want == -1), (ops, workers, sections)


def test_block_numbers_against_reference(oracle, sip):
    cases = list(rc.block_number_cases())
    assert len(cases) == len(GOLD["block_number"])
    for (nseg, lower, idx), want in zip(cases, GOLD["block_number"]):
        assert oracle.block_number(nseg, lower, idx) == want
        assert oracle.block_num2id(nseg, lower, want) == idx
        if all(lo == 1 for lo in lower):      # the product's arrays number their segments from 1
            assert sip.layout_block_number(nseg, idx) == want


def test_transposes_against_reference(oracle):
    cases = list(rc.transpose_cases())
    assert len(cases) == len(GOLD["transpose"])
    for k, (ext, perm) in enumerate(cases):
        a = rc.seeded(ext, k)
        got = oracle.block_copy(a, [1] + [p + 1 for p in perm])
        assert rc.digest(got) == GOLD["transpose"][k], (ext, perm)
        inv = np.argsort(perm)                # numpy: axes[j] = source dimension that lands at position j
        assert np.array_equal(got, np.transpose(a, inv))


def test_slices_against_reference(oracle):
    for k, (t_ext, s_ext, off) in enumerate(rc.slice_cases()):
        t = rc.seeded(t_ext, 1000 + k)
        s, ierr = oracle.block_slice(t, s_ext, off)
        assert ierr == 0 and rc.digest(s) == GOLD["slice"][k][0]
        t2, ierr = oracle.block_insert(rc.seeded(t_ext, 2000 + k), s, off)
        assert ierr == 0 and rc.digest(t2) == GOLD["slice"][k][1]


def _oracle_elementwise(oracle, op, d, s, x):
    import ctypes as C

    lib, d = oracle.lib(), np.array(d, order="F", copy=True)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    n = C.c_longlong(d.size)
    if op == "fill":
        lib.oracle_block_fill(dp(d), n, C.c_double(x))
    elif op == "scale":
        lib.oracle_block_scale(dp(d), n, C.c_double(x))
    elif op == "scale_and_copy":
        lib.oracle_block_scale_and_copy(dp(d), dp(s), n, C.c_double(x))
    elif op == "copy_data":
        d[...] = s
    elif op == "increment":
        lib.oracle_block_increment(dp(d), n, C.c_double(x))
    else:
        lib.oracle_block_accumulate(dp(d), dp(s), n)
    return d


def test_elementwise_loops_against_reference(oracle):
    for k, (op, ext, x) in enumerate(rc.elementwise_cases()):
        d, s = rc.seeded(ext, 3000 + k), rc.seeded(ext, 4000 + k)
        assert rc.digest(_oracle_elementwise(oracle, op, d, s, x)) == GOLD["elementwise"][k], (op, ext)


def test_dat_segment_tables_against_reference_setup_reader():
    """the segment tables SetupReader (reference code) decodes from the shipped inputs == tests/golden/dat_segments.json
    (decoded by aces4_b200/setup_reader.py) == SURVEY 8(d)"""
    mine = json.load(open(os.path.join(HERE, "golden", "dat_segments.json")))
    names = {"1001": "ao", "1002": "mo", "1003": "moa", "1004": "mob"}
    assert len(GOLD["dat"]) >= 5
    for fname, entry in mine.items():
        want = GOLD["dat"][fname]
        for t, ext in want["segments"].items():
            assert entry["segments"][names[t]] == ext, (fname, t)
        seg, ints = want["segments"]["1003"], want["ints"]
        assert entry["occ"] == seg[ints["baocc"] - 1: ints["eaocc"]]
        assert entry["virt"] == seg[ints["bavirt"] - 1: ints["eavirt"]]
    assert GOLD["dat"]["lccd_test.dat"]["segments"]["1001"] == [13]
    assert GOLD["dat"]["ccsdpt_test.dat"]["segments"]["1003"] == [5, 9]
    assert GOLD["dat"]["second_ccsdpt_test.dat"]["segments"]["1003"] == [5, 6]
    assert GOLD["dat"]["lccd_frozencore_test.dat"]["segments"]["1003"] == [1, 4, 8]


def test_product_checkpoint_bytes_equal_the_reference_writer(sip, tmp_path):
    """scalars-only checkpoint (arrays need a device: tests/test_gpu_persist.py) == bytes of setup::BinaryOutputFile"""
    for k, v in rc.CHECKPOINT_SCALARS.items():
        sip.persist_scalar(k, v)
    path = tmp_path / "w.ckpt"
    sip.persist_checkpoint(path)
    for k in rc.CHECKPOINT_SCALARS:
        sip.restore_scalar(k)
    assert path.read_bytes().hex() == GOLD["checkpoint_scalars_only_hex"]


def test_reference_written_checkpoint_is_what_the_format_restatement_says():
    """tests/test_persist_cpu.py::ref_checkpoint (struct-level restatement) == the reference writer's bytes, arrays included"""
    from test_persist_cpu import ref_checkpoint

    assert ref_checkpoint(rc.CHECKPOINT_SCALARS, rc.CHECKPOINT_ARRAYS).hex() == GOLD["checkpoint_hex"]
    assert ref_checkpoint(rc.CHECKPOINT_SCALARS).hex() == GOLD["checkpoint_scalars_only_hex"]


# ---------------------------------------------------------------------------------------------------------------
# live layer
# ---------------------------------------------------------------------------------------------------------------
@live
def test_fixture_is_current_with_the_live_reference():
    """regenerating a sample of every section from the live library reproduces the committed fixture"""
    for n in (1, 2, 3):
        got = "".join(rc.VERDICT_CHARS[1 + REF.block_consistency(o, w, s)] for o, w, s in rc.consistency_exhaustive(n))
        assert got == GOLD["consistency_exhaustive"][str(n)]
    assert [REF.block_consistency(o, w, s) for o, w, s in rc.consistency_sectioned()] == GOLD["consistency_sectioned"]
    assert [REF.block_number(*c)[0] for c in rc.block_number_cases()] == GOLD["block_number"]
    assert [REF.block_id_compare(*c) for c in rc.block_id_cases()] == GOLD["block_id_compare"]
    for k, (ext, perm) in enumerate(rc.transpose_cases()):
        assert rc.digest(REF.transpose_copy(rc.seeded(ext, k), perm)) == GOLD["transpose"][k]


@live
def test_random_long_access_sequences_oracle_vs_live_reference(oracle):
    for ops, workers, sections in rc.consistency_sectioned(count=1500, seed=77):
        assert oracle.block_consistency(ops, workers, sections) == REF.block_consistency(ops, workers, sections)


@live
def test_block_id_order_is_lexicographic_on_array_then_indices():
    """the reference's block-map key order (block_id.cpp) -- what `block-number order` of a slab has to agree with for a
    fixed array: comparing index tuples first-index-major"""
    for (aa, ia, ab, ib), want in zip(rc.block_id_cases(), GOLD["block_id_compare"]):
        assert REF.block_id_compare(aa, ia, ab, ib) == want
        key_a, key_b = (aa, tuple(ia)), (ab, tuple(ib))
        assert want == (0 if key_a == key_b else -1 if key_a < key_b else 1)


@live
def test_sial_known_answer_transposes_through_reference_block(oracle):
    """BasicSial.transpose_tmp / transpose4d_tmp (test_basic_sial.cpp:653-693,1285-1327) through sip::Block::transpose_copy"""
    a = oracle.fill_sequential((8, 8, 8), 53.0)
    b = REF.transpose_copy(a, [2, 0, 1])                       # b[j,k,i] = a[i,j,k]
    for i, j, k in ((0, 0, 0), (1, 2, 3), (7, 6, 5)):
        assert b[j, k, i] == a[i, j, k]
    assert np.array_equal(b, oracle.permute_labels([2, 3, 1], [1, 2, 3], a))
    a4 = oracle.fill_sequential((5, 5, 5, 1), 53.0)
    b4 = REF.transpose_copy(a4, [2, 1, 0, 3])                  # b[k,j,i,l] = a[i,j,k,l]
    assert np.array_equal(b4, np.transpose(a4, (2, 1, 0, 3)))
    assert np.array_equal(b4, oracle.permute_labels([3, 2, 1, 4], [1, 2, 3, 4], a4))


@live
def test_sum_op_and_scale_through_reference_block(oracle):
    """BasicSial.sum_op (d = a + c; e = d - c, 20x20 sequential from 100 / 50) with the reference's accumulate loop"""
    a, c = oracle.fill_sequential((20, 20), 100.0), oracle.fill_sequential((20, 20), 50.0)
    d = REF.block_op(REF.ACCUMULATE, REF.block_op(REF.COPY_DATA, np.zeros((20, 20), order="F"), a), c)
    assert np.array_equal(d, a + c)
    e = REF.block_op(REF.ACCUMULATE, d, REF.block_op(REF.SCALE, c, x=-1.0))
    assert np.array_equal(e, a)


@live
def test_dat_files_product_reader_vs_live_setup_reader():
    from aces4_b200.setup_reader import read_setup

    names = {"ao": 1001, "mo": 1002, "moa": 1003, "mob": 1004}
    n = 0
    for fname in GOLD["dat"]:
        path = os.path.join(REF.REFERENCE_ROOT, "test", fname)
        if not os.path.exists(path):
            continue
        mine = read_setup(open(path, "rb").read())
        assert mine["trailing_bytes"] == 0
        for kind, ext in mine["segments"].items():
            assert REF.setup_segments(path, names[kind]) == ext, (fname, kind)
        for key in list(mine["ints"])[:12]:
            assert REF.setup_predefined_int(path, key) == mine["ints"][key]
        for key in list(mine["scalars"])[:6]:
            assert REF.setup_predefined_scalar(path, key) == mine["scalars"][key]
        n += 1
    assert n >= 5


@live
def test_checkpoint_files_cross_read(sip, tmp_path):
    """product-written checkpoint parsed by the reference's BinaryInputFile; reference-written one restored by the product"""
    scal = {"e_scf": -75.58432674274046, "iter": 12.0}
    for k, v in scal.items():
        sip.persist_scalar(k, v)
    mine = tmp_path / "mine.ckpt"
    sip.persist_checkpoint(mine)
    for k in scal:
        sip.restore_scalar(k)
    kinds = [REF.K_INT] + [REF.K_STRING, REF.K_DOUBLE] * len(scal) + [REF.K_INT]
    recs = REF.stream_read(str(mine), kinds)
    assert recs[0] == (REF.K_INT, 2) and recs[-1] == (REF.K_INT, 0)
    assert {recs[1][1]: recs[2][1], recs[3][1]: recs[4][1]} == scal

    theirs = tmp_path / "theirs.ckpt"
    REF.stream_write(str(theirs), [(REF.K_INT, 2), (REF.K_STRING, "alpha  "), (REF.K_DOUBLE, 0.5), (REF.K_STRING, "beta"),
                                   (REF.K_DOUBLE, -2.25), (REF.K_INT, 0)])
    sip.persist_init_from_checkpoint(theirs)
    assert sip.restore_scalar("alpha") == 0.5 and sip.restore_scalar("beta") == -2.25
    assert struct.unpack("<i", theirs.read_bytes()[:4])[0] == 2
