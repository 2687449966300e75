"""The reference's block-operation unit tests (test/test_basic_sial.cpp, test/test_sial.cpp: the known-answer tests of SURVEY.md
section 8c) driven from THEIR OWN SIAL programs -- tests/golden/ref_unit_programs/*.sialx = src/sialx/test/*.sialx verbatim
(scripts/make_unit_program_goldens.py) -- through the SIAL front-end on a backend, with the segment tables and constants each C++
test sets up and the assertions it makes (restated from the C++ loops).  Shared by the CPU (oracle backend) and GPU (libsipgpu)
test files; `make_backend(arrays)` builds the backend, `to_numpy(handle)` reads a block back."""
import itertools
import os

import numpy as np

from aces4_b200.sial_frontend import Program, Walker
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def text(name):
    return open(os.path.join(HERE, "golden", "ref_unit_programs", name + ".sialx")).read()


def _fill(kind):
    """the reference's test super-instructions fill_block_cyclic / fill_block_sequential (super_instructions/.../fill_block_*.f):
    element n (column-major, 0-based) = ((n + start - 1) mod 20) + 1  /  start + n"""
    def si(w, args, bare):
        if not args:      # `execute fill_block_sequential a 1.0`: a whole static array, contiguous in the SIP (static_array_test.sialx)
            name, start = bare[0], float(bare[1])
            _, dims = w._static_blocks(name)
            w._static_scatter(name, (oracle.fill_cyclic if kind == "cyclic" else oracle.fill_sequential)(tuple(dims), start))
            return
        name, labs = args[0]
        start = float(bare[0]) if bare[0].replace(".", "", 1).replace("-", "", 1).isdigit() else w.be.value(w.scalars[bare[0]])
        h = w._write(name, labs)
        shape = w._shape(labs)
        w.be.set_from_host(h, (oracle.fill_cyclic if kind == "cyclic" else oracle.fill_sequential)(shape, start))
    return si


EXTRA_SI = {"fill_block_cyclic": _fill("cyclic"), "fill_block_sequential": _fill("sequential"),
            "list_blocks_with_number": lambda w, a, b: None, "list_block_map": lambda w, a, b: None,
            "one_arg_no_op": lambda w, a, b: None, "enable_all_rank_print": lambda w, a, b: None,
            "disable_all_rank_print": lambda w, a, b: None}


def dim_segments(prog, label, seg_tables, constants):
    """segment extents of the range a declared index `label` runs over (a simple index: as many one-element segments as values)"""
    kind = prog.index_kind[label]
    val = lambda x: int(x) if str(x).isdigit() else int(constants[x])      # noqa: E731
    if kind == "s":
        lo, hi = prog.simple_range[label]
        return [1] * (val(hi) - val(lo) + 1)
    if ":" in kind:
        typ, lo, hi = kind.split(":")
        return list(seg_tables[typ][val(lo) - 1: val(hi)])
    if kind == "ao":
        return list(seg_tables["ao"][: val("norb")] if "norb" in constants else seg_tables["ao"])
    lo, hi = {"o": ("baocc", "eaocc"), "v": ("bavirt", "eavirt"), "p": ("baocc", "eavirt"), "pa": (1, "eavirt")}[kind]
    return list(seg_tables["mo"][val(lo) - 1: val(hi)])


def expected_blocks(name):
    """the blocks the reference's expected-output fixture of a test holds (test/expected_output/<name>.txt, copied verbatim into
    tests/golden/ref_expected_output/): [(array, index values, values in memory order)] in the order printed"""
    import re
    out, cur = [], None
    for line in open(os.path.join(HERE, "golden", "ref_expected_output", name + ".txt")):
        m = re.match(r"\s*\d+:\s+printing (\d+) of (\d+) elements of (block|contiguous array) (\w+)(?:\[([\d,\s]*)\])? in the order stored", line)
        if m:
            cur = (m.group(4), tuple(int(x) for x in m.group(5).split(",")) if m.group(5) else (), [])
            out.append(cur)
            assert m.group(1) == m.group(2)
            want = int(m.group(1))
            continue
        if cur is not None and len(cur[2]) < want:
            try:
                cur[2].extend(float(x) for x in line.split())
            except ValueError:
                cur = None
    assert all(len(v) > 0 for _, _, v in out)
    return out


def run(name, make_backend, ao=None, mo=None, constants=None, arrays=(), backend=None, rank=0, world=1, print_hook=None):
    prog = Program(text(name))
    be = backend or make_backend(prog, {"ao": ao or [], "mo": mo or []}, constants or {})
    segs = {"ao": list(ao or []), "pa": list(mo or [])}
    c = dict(constants or {})
    if mo and "baocc" in c:
        segs["o"] = mo[c["baocc"] - 1: c["eaocc"]]
        segs["v"] = mo[c["bavirt"] - 1: c["eavirt"]]
    w = Walker(prog, be, segs, rank=rank, world=world, constants=c, seg_tables={"ao": ao or [], "mo": mo or []}, extra_si=EXTRA_SI,
               index_base={"o": c.get("baocc", 1) - 1, "v": c.get("bavirt", 1) - 1})
    w.print_hook = print_hook
    w.run()
    return w, be


# ---- the tests (each returns nothing, asserts) ----------------------------------------------------------------------
def contraction_small_test(make_backend, to_numpy):
    """BasicSial.contraction_small_test (test_basic_sial.cpp:695-770): c[i,l] = a[i,j,k,l]*b[j,k], segment 15"""
    w, _ = run("contraction_small_test", make_backend, ao=[15, 15, 15, 15, 15, 15, 15, 15, 14, 14, 14, 14, 12, 12])
    c_data = to_numpy(w.block_of("c", (1, 1))).ravel(order="F")
    n = 15
    a = ((np.arange(n ** 4) % 20) + 1).astype(float).reshape(n, n, n, n)      # the C++ test's row-major arrays
    b = ((np.arange(n ** 2) % 20) + 1).astype(float).reshape(n, n)
    for i in range(n):
        for l in range(n):
            assert np.sum(a[i, :, :, l] * b) == c_data[i * n + l]


def contraction_small_test2(make_backend, to_numpy):
    """BasicSial.contraction_small_test2 (:773-815) + test/test_contraction_small2.F: c[mu,i1,a1,i] = b[lambda,a1]*a[mu,i1,i,lambda];
    ao 9, moa {5, 4} (occupied = segment 1, virtual = segment 2)"""
    w, _ = run("contraction_small_test2", make_backend, ao=[9], mo=[5, 4], constants={"baocc": 1, "eaocc": 1, "bavirt": 2, "eavirt": 2})
    c = to_numpy(w.block_of("c", (1, 1, 1, 1)))
    a = oracle.fill_cyclic((9, 5, 5, 9), 1.0)
    b = oracle.fill_cyclic((9, 4), 1.0)
    assert np.array_equal(c, np.einsum("mxyq,qz->mxzy", a, b))


def transpose_tmp(make_backend, to_numpy):
    """BasicSial.transpose_tmp (:653-693) + test_transpose_op.F: b[j,k,i] = a[i,j,k], 8 x 8 x 8, sequential from 53"""
    w, _ = run("transpose_tmp", make_backend, ao=[8, 12, 10], constants={"norb": 3})
    b = to_numpy(w.block_of("b", (1, 1, 1)))
    a = oracle.fill_sequential((8, 8, 8), 53.0)
    assert np.array_equal(b, np.transpose(a, (1, 2, 0)))


def transpose4d_tmp(make_backend, to_numpy):
    """BasicSial.transpose4d_tmp (:1285-1327) + test_transpose4d_op.F: b[k,j,i,l] = a[i,j,k,l]; moa {1, 5, 4}: occupied = segment 1
    (one orbital), virtual = segment 2 (five): the block the test reads is b[2,1,2,1]"""
    w, _ = run("transpose4d_tmp", make_backend, mo=[1, 5, 4], constants={"norb": 3, "baocc": 1, "eaocc": 1, "bavirt": 2, "eavirt": 2})
    b = to_numpy(w.block_of("b", (1, 1, 1, 1)))          # (first virtual, first occupied, ...) = absolute segments (2, 1, 2, 1)
    a = oracle.fill_sequential((5, 1, 5, 1), 53.0)
    assert b.shape == (5, 1, 5, 1) and np.array_equal(b, np.transpose(a, (2, 1, 0, 3)))


def transpose4d_square_tmp(make_backend, to_numpy):
    """BasicSial.transpose4d_square_tmp (:1329-1406): 8^4 cyclic; esum1 = a*a, esum2 = b*b, esum3 = a[i,j,k,l]*b[k,j,i,l]"""
    w, be = run("transpose4d_square_tmp", make_backend, ao=[8, 8, 8], constants={"norb": 3})
    for k in ("esum1", "esum2", "esum3"):
        assert be.value(w.scalars[k]) == 204 * 2870 + 1496 == 586976.0


def contract_to_scalar(make_backend, to_numpy):
    """BasicSial.contract_to_scalar (:1037-1084): x = a[i,j]*b[i,j], 8 x 8, a cyclic from 1, b cyclic from 5"""
    w, be = run("contract_to_scalar", make_backend, ao=[8, 8], constants={"norb": 2})
    assert be.value(w.scalars["x"]) == float(sum((((c % 20) + 1) * (((c + 4) % 20) + 1)) for c in range(64)))


def sum_op(make_backend, to_numpy):
    """BasicSial.sum_op (:817-916): d = a + c; e = d - c; 20 x 20, sequential from 100 / 50"""
    w, _ = run("sum_op_test", make_backend, ao=[20, 5], constants={"norb": 2})
    n = np.arange(400).reshape((20, 20), order="F")
    assert np.array_equal(to_numpy(w.block_of("d", (1, 1))), 150.0 + 2 * n)
    assert np.array_equal(to_numpy(w.block_of("e", (1, 1))), 100.0 + n)


def self_multiply_test(make_backend, to_numpy):
    """BasicSial.self_multiply_test (:1111-1160): a sequential from 100; a += a; a *= 1.5  ->  3 (100 + n)"""
    w, _ = run("self_multiply_test", make_backend, ao=[20, 5], constants={"norb": 2})
    n = np.arange(400).reshape((20, 20), order="F")
    assert np.array_equal(to_numpy(w.block_of("a", (1, 1))), 3.0 * (100.0 + n))


def put_test(make_backend, to_numpy):
    """Sial.put_test (test_sial.cpp:282-318): put a[i,j] = k; get; x = a*a -> result[k] = k^2 seg_i seg_j"""
    segs = [2, 3, 2]
    w, _ = run("put_test", make_backend, ao=segs, constants={"norb": 3, "norb_squared": 9})
    for i, j in itertools.product(range(3), repeat=2):
        k = i * 3 + j + 1
        assert to_numpy(w.block_of("result", (k,))).ravel()[0] == float(k * k * segs[i] * segs[j])


def get_mpi(make_backend, to_numpy):
    """Sial.get_mpi (:485-520): put b = a; put c = 0; put c += a twice; barrier; get; a = b + c = 126 in every element"""
    segs = [2, 3, 4, 2]
    w, _ = run("get_mpi", make_backend, ao=segs, constants={"norb": 4})
    for i, j in itertools.product(range(1, 5), repeat=2):
        blk = to_numpy(w.block_of("a", (i, j)))
        assert blk.shape == (segs[i - 1], segs[j - 1]) and np.all(blk == 126.0)


def put_accumulate_stress(make_backend, to_numpy):
    """Sial.put_accumulate_stress (:1072-1113): pardo k = 1..20: put c[i,j] += a, += aa, += a, += aa with a = i, aa = j:
    every element of c[i,j] = 20 (2 i + 2 j) -- many-writer accumulate correctness"""
    segs = [2, 3, 2, 2]
    w, _ = run("put_accumulate_stress", make_backend, ao=segs, constants={"norb": 4, "kmax": 20})
    for i, j in itertools.product(range(1, 5), repeat=2):
        blk = to_numpy(w.block_of("a", (i, j)))
        assert blk.shape == (segs[i - 1], segs[j - 1]) and np.all(blk == 20.0 * (2 * i + 2 * j))


def gpu_path_programs(make_backend, to_numpy):
    """the programs of the reference's dormant CUDA path (src/sialx/test/gpu_*.sialx: the same operations between gpu_on / gpu_put /
    gpu_allocate / gpu_get / gpu_free / gpu_off): the answers of their host twins"""
    w, _ = run("gpu_contraction_small_test", make_backend, ao=[15, 15, 15, 15, 15, 15, 15, 15, 14, 14, 14, 14, 12, 12])
    a, b = oracle.fill_cyclic((15,) * 4, 1.0), oracle.fill_cyclic((15, 15), 1.0)
    assert np.array_equal(to_numpy(w.block_of("c", (1, 1))), np.einsum("ijkl,jk->il", a, b))
    w, _ = run("gpu_transpose_tmp", make_backend, ao=[8, 12, 10], constants={"norb": 3})
    assert np.array_equal(to_numpy(w.block_of("b", (1, 1, 1))), np.transpose(oracle.fill_sequential((8, 8, 8), 53.0), (1, 2, 0)))
    w, be = run("gpu_contract_to_scalar", make_backend, ao=[8, 8], constants={"norb": 2})
    assert be.value(w.scalars["x"]) == float(sum((((c % 20) + 1) * (((c + 4) % 20) + 1)) for c in range(64)))
    n = np.arange(400).reshape((20, 20), order="F")
    w, _ = run("gpu_self_multiply_test", make_backend, ao=[20, 5], constants={"norb": 2})
    assert np.array_equal(to_numpy(w.block_of("a", (1, 1))), 3.0 * (100.0 + n))
    w, _ = run("gpu_sum_op_test", make_backend, ao=[20, 5], constants={"norb": 2})
    assert np.array_equal(to_numpy(w.block_of("a", (1, 1))), 100.0 + n) and np.array_equal(to_numpy(w.block_of("c", (1, 1))), 50.0 + n)
    run("gpu_ops", make_backend, ao=[3, 4])


def put_initialize_and_increment(make_backend, to_numpy):
    """Sial.put_initialize / put_increment (test_sial.cpp:322-420): `put a[i,j] = x`, `put a[i,j] += x`, `put a[i,j] *= -1.0` with a scalar
    (SialOpsParallel::put_initialize / put_increment / put_scale): result[k] = k^2 seg_i seg_j; every element of a[i,j] = -k"""
    segs = [2, 3, 2]
    w, _ = run("put_initialize", make_backend, ao=segs, constants={"norb": 3, "norb_squared": 9})
    for i, j in itertools.product(range(3), repeat=2):
        k = i * 3 + j + 1
        assert to_numpy(w.block_of("result", (k,))).ravel()[0] == float(k * k * segs[i] * segs[j])
    w, be = run("put_increment", make_backend, ao=segs, constants={"norb": 3, "norb_squared": 9})
    for i, j in itertools.product(range(3), repeat=2):
        blk = to_numpy(be.array_block("a", (i + 1, j + 1), (segs[i], segs[j])))
        assert np.all(blk == -float(i * 3 + j + 1))


def persistence_between_programs(make_backend, to_numpy):
    """Sial.persistent_distributed_array_mpi (test_sial.cpp:640-700): program 1 fills b (put) and c (put += twice) and persists them, program 2
    restores them: a = b + c = 3 x (sequential values from (i-1) norb + j); BasicSial.persistent_static_array_test and
    persistent_scalars (test_basic_sial.cpp): a static array the first program filled and a scalar, restored by the second"""
    segs = [2, 3]
    run("persistent_distributed_array_mpi1", make_backend, ao=segs, constants={"norb": 2})
    w, _ = run("persistent_distributed_array_mpi2", make_backend, ao=segs, constants={"norb": 2})
    for i, j in itertools.product((1, 2), repeat=2):
        blk = to_numpy(w.block_of("a", (i, j)))
        first = (i - 1) * 2 + j
        assert np.array_equal(blk.ravel(order="F"), 3.0 * (first + np.arange(blk.size)))
        assert np.array_equal(to_numpy(w.block_of("lb", (i, j))).ravel(order="F"), 1.0 * (first + np.arange(blk.size)))
    # Sial.persistent_distributed_array_n_of_three (test_sial.cpp:822-900): the SECOND and the THIRD program both restore "savedb" /
    # "savedc" although the second does not persist them again -- the servers' files outlive a restore
    # (disk_backed_block_map.restore_persistent_array); the harness hands the arrays over again under the same labels
    def a_is_three_times_the_sequence(w):
        for i, j in itertools.product((1, 2), repeat=2):
            blk = to_numpy(w.block_of("a", (i, j)))
            assert np.array_equal(blk.ravel(order="F"), 3.0 * ((i - 1) * 2 + j + np.arange(blk.size)))
    run("persistent_distributed_array_one_of_three", make_backend, ao=segs, constants={"norb": 2})
    w, be = run("persistent_distributed_array_two_of_three", make_backend, ao=segs, constants={"norb": 2})
    a_is_three_times_the_sequence(w)
    for name, label in (("b", "savedb"), ("c", "savedc")):
        if hasattr(be.arrays[name], "persist"):
            be.arrays[name].persist(label)          # libsipgpu: sipgpu_array_persist
        else:
            type(be).registry[label] = be.arrays[name]
    w, be = run("persistent_distributed_array_three_of_three", make_backend, ao=segs, constants={"norb": 2})
    a_is_three_times_the_sequence(w)
    run("persistent_static_array_test1", make_backend, ao=segs, constants={"norb": 2})
    w, _ = run("persistent_static_array_test2", make_backend, ao=segs, constants={"norb": 2})
    for i, j in itertools.product((1, 2), repeat=2):
        blk = to_numpy(w.block_of("lb", (i, j)))
        assert np.array_equal(blk.ravel(order="F"), 1.0 + np.arange(blk.size))
    run("persistent_scalars_1", make_backend, constants={"x": 3.456, "y": -0.1})
    w, be = run("persistent_scalars_2", make_backend, constants={"y": -0.1})
    assert be.value(w.scalars["e"]) == 6.0 and abs(be.value(w.scalars["x"]) - 4.456) < 1e-15


# ---- host-only programs (the pardo work distribution and the interpreter's arithmetic: no block operation is issued) ----
def pardo_loops(make_backend, to_numpy):
    """Sial.pardo_loop / pardo_loop_corner_case / pardo_loop_with_pragma (test_sial.cpp:82-216): pardo_loop_<n>d counts its
    iterations per worker and sums the counters with `collective total += (scalar)counter`: total == prod(upper - lower + 1)
    for every number of workers (the work distribution of BalancedTaskAllocParallelPardoLoop::do_update: every iteration on
    exactly one worker).  The collective is the sum over the emulated ranks."""
    for lower, upper in (([3, 2, 4, 1, 99, -1], [7, 6, 5, 1, 101, 2]), ([1] * 6, [1] * 6)):
        for nd in range(1, 7):
            consts = {f"lower{i}": lower[i] for i in range(nd)} | {f"upper{i}": upper[i] for i in range(nd)}
            want = int(np.prod([upper[i] - lower[i] + 1 for i in range(nd)]))
            for world in (1, 2, 3, 7):
                total = 0
                for rank in range(world):
                    w, be = run(f"pardo_loop_{nd}d", make_backend, constants=consts, rank=rank, world=world)
                    assert be.value(w.scalars["total"]) == be.value(w.scalars["counter"])      # one rank's share
                    total += int(be.value(w.scalars["counter"]))
                assert total == want, (nd, world, total, want)
    for world in (1, 3):
        counts = []
        for r in range(world):
            w, be = run("pardo_loop_with_pragma", make_backend, rank=r, world=world,
                        constants={"lower0": 3, "upper0": 7, "lower1": 2, "upper1": 6, "lower2": 4, "upper2": 5})
            counts.append(int(be.value(w.scalars["counter"])))
        assert sum(counts) == 5 * 5 * 2
    w, be = run("pardo_loop", make_backend)
    assert int(be.value(w.scalars["counter"])) == 5 * 4 * 4
    # pardo_with_where (test_sial.cpp:1039-1070: runs to completion) -- here also: the where clauses leave the
    # mu < nu, lambda < sigma, mu < lambda, all-different quadruples, each on exactly one of 3 workers
    seen = []
    for r in range(3):
        w, _ = run("pardo_with_where", make_backend, ao=[2, 3, 2, 2], constants={"norb": 4}, rank=r, world=3)
        seen.append(w.iteration)
    assert len(set(seen)) == 1     # every worker counts every where-true iteration of the section


def interpreter_arithmetic(make_backend, to_numpy):
    """BasicSial.scalar_ops, int_ops, int_self_ops, ifelse, index_scalar_cast (test_basic_sial.cpp:285-296, 451-524, 1477-1499)"""
    w, be = run("scalar_ops", make_backend)
    val = lambda n: be.value(w.scalars[n])      # noqa: E731
    for name, want in (("l", 42.0), ("nl", -42.0), ("s0", 0.0), ("si0", 0.0), ("sd", 21.0), ("sr0", 16), ("sr1", 4), ("e0", 16),
                       ("ci0", 4), ("ci1", 16), ("ci2", -28), ("re1", 1), ("re2", -1), ("rgt2", 2), ("rgt3", 15), ("rgt4", 15),
                       ("rgt5", 10), ("rgt6", 10), ("rgt7", 10), ("rgt8", 10), ("rgt9", 10)):
        assert val(name) == want, (name, val(name), want)      # ci0 = (int) 3.75 = 4: the SIP's cast is lrint (sial_math.cpp:39)
    w, be = run("int_ops", make_backend)
    for name, want in dict(l=42, nl=-42, s0=0, si0=0, sd=21, sr0=10, sr1=-2, e0=77, ci0=4, ci1=12, ci2=3, re1=1, re2=-1, rgt2=2, rgt3=15,
                           rgt4=15, rgt5=10, rgt6=10, rgt7=10, rgt8=15, rgt9=10).items():
        assert be.value(w.scalars[name]) == want, (name, be.value(w.scalars[name]), want)      # int / int truncates
    w, be = run("ifelse", make_backend)
    assert be.value(w.scalars["eq_counter"]) == 4 and be.value(w.scalars["neq_counter"]) == 20
    w, be = run("int_self_ops", make_backend)
    assert [be.value(w.scalars[n]) for n in "xyzw"] == [76, 44, -28, 15200]
    w, be = run("exit_statement_test", make_backend, ao=[2, 3, 4] * 5, constants={"norb": 3})      # :623-650: `exit` leaves ONE loop
    assert be.value(w.scalars["counter_j"]) == 12 and be.value(w.scalars["counter_i"]) == 4
    w, be = run("return_sval_test", make_backend)      # :1469-1475 (runs): `x = a[i,j]` and `execute return_sval a[i,j] y` agree
    assert be.value(w.scalars["x"]) == be.value(w.scalars["y"]) == 4 * 4 + 3
    w, be = run("index_scalar_cast", make_backend, ao=[2, 2, 1, 3], constants={"norb": 4})
    assert be.value(w.scalars["count"]) == 4 and be.value(w.scalars["count2"]) == 1


HOST_ONLY = (pardo_loops, interpreter_arithmetic)


def printed_blocks_equal_the_reference_fixtures(make_backend, to_numpy):
    """BasicSial.static_array_test (blocks extracted from a static array that was filled as ONE contiguous array, test_basic_sial.cpp:918-940),
    tmp_arrays, tmp_arrays_2 (:526-579), local_arrays, local_arrays_wild (test_simple.cpp:785-830): the reference compares what the program PRINTS with a fixture file
    (`EXPECT_EQ(controller.expectedOutput(), output.str())`); here every printed block is compared, in order, with the blocks of that
    same fixture"""
    for name, ao, consts in (("static_array_test", [3, 4], {"norb": 2, "x": 3.456}), ("tmp_arrays", [2, 3, 4], {"norb": 3, "x": 3.456}),
                             ("tmp_arrays_2", [2, 3, 4], {"norb": 3, "x": 3.456}), ("local_arrays", [2, 3], {"norb": 2, "x": 3.456}),
                             ("local_arrays_wild", [2, 3], {"norb": 2, "x": 3.456})):      # `allocate a[i,*]` row by row
        printed = []
        w, be = run(name, make_backend, ao=ao, constants=consts,
                    print_hook=lambda arr, idx, a: printed.append((arr, idx, np.asarray(a).ravel(order="F").copy())))
        want = [b for b in expected_blocks(name) if b[1]]
        assert len(printed) == len(want) > 0, (name, len(printed), len(want))
        for (arr, idx, vals), (warr, widx, wvals) in zip(printed, want):
            assert (arr, idx) == (warr, widx) and np.array_equal(vals, np.array(wvals)), (name, arr, idx)
        if name == "static_array_test":      # `print a`: the whole contiguous array, in memory order
            whole = [b for b in expected_blocks(name) if not b[1]]
            assert len(whole) == 1 and np.array_equal(w._static_dense("a").ravel(order="F"), np.array(whole[0][2]))
    # scalar_valued_blocks (test_basic_sial.cpp:581-621): one-element blocks over simple indices read and written as numbers
    printed = []
    run("scalar_valued_blocks", make_backend, print_hook=lambda arr, idx, a: printed.append(float(np.asarray(a).ravel()[0])))
    assert printed == [1.0, 2.0, 3.0] + [float(j + k) for k in (1, 2, 3) for j in (1, 2, 3, 4, 5)]
    # cast_indices_to_simple (:1428-1466): `s = a[(index)j]` -- the value of a segment index addresses a simple-index dimension
    w, be = run("cast_indices_to_simple", make_backend, ao=[2, 2, 5, 5, 5], constants={"norb": 5})
    for j, ext in enumerate([2, 2, 5, 5, 5], start=1):
        assert np.array_equal(to_numpy(w.block_of("b", (j,))).ravel(), np.full(ext, 5.0 - j))
    # simple_indices_assignments (:1086-1109)
    w, be = run("simple_indices_assignments", make_backend, ao=[8, 8], constants={"norb": 2, "x": 3.456})
    assert be.value(w.scalars["x"]) == 50 and be.value(w.scalars["y"]) == 50


def eom_idiom_contiguous_local(make_backend, to_numpy):
    """Sial.contig_local3 (test_sial.cpp:667-697; the reference only runs it -- it is the regression test of a crash at "line 6289" of
    the EOM program): rank-5 served arrays with a leading simple index, prepared / requested block by block and staged through a
    contiguous local array `CLRB2_aa[1:eom_subspc, a:a, i:i, a1:a1, i1:i1]`.  Closed form of what it prints: slot ksub of every
    (a,i,a1,i1) block holds (scalar)ksub for ksub <= eom_roots and the zeros of the fresh allocation above"""
    printed = []
    run("contig_local3", make_backend, mo=[2, 3, 4, 1, 4, 4, 4, 4],
        constants=dict(eom_roots=4, eom_subspc=8, baocc=1, eaocc=3, bavirt=4, eavirt=8, norb=8),
        print_hook=lambda n, idx, a: printed.append((idx[0], np.asarray(a))))
    assert len(printed) == 8 * 5 * 3 * 5 * 3
    for ksub, a in printed:
        assert np.all(a == (float(ksub) if ksub <= 4 else 0.0)), ksub


def runs_to_completion(make_backend, to_numpy):
    """BasicSial.tmp_arrays / tmp_arrays_2 / block_scale_assign (:526-650), Sial.put_accumulate_mpi: the reference compares printed
    output; here: the programs run to completion through the same statements (block fill / scale / add / subtract / copy with
    permutation, scalar-valued fills from index casts)"""
    for name, ao, c in (("tmp_arrays", [2, 3, 4], {"norb": 3}), ("tmp_arrays_2", [2, 3, 4], {"norb": 3}),
                        ("block_scale_assign", [2, 3, 4], {"norb": 3}), ("put_accumulate_mpi", [2, 3, 4, 2], {"norb": 4})):
        run(name, make_backend, ao=ao, constants=c)


ALL = (contraction_small_test, contraction_small_test2, transpose_tmp, transpose4d_tmp, transpose4d_square_tmp, contract_to_scalar,
       sum_op, self_multiply_test, put_test, get_mpi, put_accumulate_stress, put_initialize_and_increment, gpu_path_programs,
       persistence_between_programs, printed_blocks_equal_the_reference_fixtures, eom_idiom_contiguous_local, runs_to_completion)
