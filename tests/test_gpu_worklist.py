"""GPU parity tests of the deferred op stream (aces4_b200/csrc/worklist.cu) and the SIAL front-end on the device
backend: the SAME per-block call stream executed (a) op-at-a-time, (b) recorded and scheduled into batched launches,
(c) on the CPU oracle.  Tolerance: 1e-10 relative on blocks (BASELINE.json north_star); recorded vs op-at-a-time differ
only by the summation order inside fused chains."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-10


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init(0)
    return s.api


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def run_stream(sip, seed, record, nops=220, v=12):
    rnd = random.Random(seed)
    rng = np.random.default_rng(seed)
    p2, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    live = [sip.DeviceBlock.from_numpy(rng.uniform(-1, 1, (v, v))) for _ in range(6)]
    if record:
        sip.wl_begin()
    for _ in range(nops):
        k = rnd.randrange(10)
        pick = lambda: rnd.choice(live)  # noqa: E731
        if k == 0:
            pick().fill(0.0 if rnd.random() < 0.5 else 0.25)
        elif k == 1:
            pick().scale(0.5)
        elif k == 2:
            d, s = pick(), pick()
            if d is not s:
                d.axpy(s, rnd.choice([1.0, -1.0, 0.5]))
        elif k == 3:
            d, s = pick(), pick()
            if d is not s:
                d.scale_and_copy(s, 0.75)
        elif k == 4:
            d, a, b = pick(), pick(), pick()
            d.set_add_sub(a, b, rnd.choice([1.0, -1.0]))
        elif k == 5:
            d, s = pick(), pick()
            if d is not s:
                sip.permute(s, [1, 2, 1], out=d)
        elif k in (6, 7):
            d, a, b = pick(), pick(), pick()
            if d is not a and d is not b:
                sip.contract(p2, a, b, (v, v), out=d, alpha=0.125, beta=rnd.choice([0.0, 1.0, 1.0]))
        elif k == 8:
            a, b, d = pick(), pick(), pick()
            t = sip.DeviceBlock((v, v))
            sip.contract(p2, a, b, (v, v), out=t, alpha=0.125)
            d.axpy(t, 1.0)
            t.free()
        else:
            s, d = pick(), pick()
            t = sip.DeviceBlock((v, v))
            sip.permute(s, [1, 2, 1], out=t)
            d.axpy(t, -0.5)
            t.free()
    stats = sip.wl_end() if record else None
    out = [b.to_numpy() for b in live]
    for b in live:
        b.free()
    return out, stats


@pytest.mark.parametrize("seed", range(8))
def test_random_streams_recorded_equal_op_at_a_time(sip, seed):
    eager, _ = run_stream(sip, seed, record=False)
    rec, st = run_stream(sip, seed, record=True)
    for a, b in zip(rec, eager):
        assert np.all(np.isfinite(b))
        assert rel(a, b) <= 1e-12
    assert st["scheduled"] < st["recorded"] and st["fused_accumulates"] > 0


def test_implicit_flush_on_blocking_reads(sip):
    a = sip.DeviceBlock.from_numpy(np.arange(24.0).reshape(4, 6))
    with sip.recording():
        a.scale(2.0)
        before = sip.kernel_launches()
        got = a.to_numpy()          # d2h is blocking: the recording is drained first
        assert sip.kernel_launches() > before
        assert np.array_equal(got, 2.0 * np.arange(24.0).reshape(4, 6))
        a.increment(1.0)
        assert abs(a.norm2() - float(np.sum((2.0 * np.arange(24.0) + 1.0) ** 2))) < 1e-9
    assert sip.lib().sipgpu_wl_recording() == 0


def test_hhladder_chain_matches_oracle(sip, oracle):
    """do i1, j1: T = T2old[a,i1,b,j1]*V[i,i1,j,j1]; D += T (rlccd_rhf.sialx:342-355), one destination, 3x3 segments"""
    v, o, n = 10, 6, 3
    rng = np.random.default_rng(5)
    T2 = [[rng.uniform(-1, 1, (v, o, v, o)) for _ in range(n)] for _ in range(n)]
    V = [[rng.uniform(-1, 1, (o, o, o, o)) for _ in range(n)] for _ in range(n)]
    dlab, llab, rlab = [1, 2, 3, 4], [1, 5, 3, 6], [2, 5, 4, 6]
    want = np.zeros((v, o, v, o), order="F")
    for i1 in range(n):
        for j1 in range(n):
            t, ierr = oracle.contract_labels(dlab, [v, o, v, o], llab, T2[i1][j1], rlab, V[i1][j1])
            assert ierr == 0
            want += t
    dT2 = [[sip.DeviceBlock.from_numpy(x) for x in row] for row in T2]
    dV = [[sip.DeviceBlock.from_numpy(x) for x in row] for row in V]
    D = sip.DeviceBlock((v, o, v, o))
    launches0 = sip.kernel_launches()
    with sip.recording() as rec:
        D.fill(0.0)
        for i1 in range(n):
            for j1 in range(n):
                t = sip.DeviceBlock((v, o, v, o))
                sip.contract_labels(dlab, (v, o, v, o), llab, dT2[i1][j1], rlab, dV[i1][j1], out=t)
                D.accumulate(t)
                t.free()
    assert rec.stats["chains"] == 1 and rec.stats["chain_pairs"] == n * n and rec.stats["scheduled"] == 1
    assert sip.kernel_launches() - launches0 == 1        # 19 recorded ops -> ONE kernel launch
    assert rel(D.to_numpy(), want) <= TOL


def test_put_accumulate_stress_recorded(sip):
    """Sial.put_accumulate_stress (test_sial.cpp:1072-1113): pardo k(1..20): put c[i,j] += a; += aa; += a; += aa with
    a = i, aa = j  ->  every element of c[i,j] = 20*(2i+2j); here through the recorded stream (red.add commutes)."""
    segs = [[2, 3, 2], [2, 3, 2]]
    c = sip.DistArray(segs)
    c.fill_local(0.0)
    sip.sync()
    with sip.recording() as rec:
        for k in range(20):
            for i in range(1, 4):
                for j in range(1, 4):
                    shape = c.block_shape((i, j))
                    a, aa = sip.DeviceBlock(shape), sip.DeviceBlock(shape)
                    a.fill(float(i))
                    aa.fill(float(j))
                    for blk in (a, aa, a, aa):
                        c.put_accumulate((i, j), blk)
                    a.free()
                    aa.free()
    assert rec.stats["levels"] == 2 and rec.stats["launches"] == 2     # all fills, then all red.adds
    for i in range(1, 4):
        for j in range(1, 4):
            got = c.get((i, j)).to_numpy()
            assert np.all(got == 20.0 * (2 * i + 2 * j))
    c.destroy()


def test_opaque_ops_keep_program_order(sip, oracle):
    """slices of a static array, a scalar contraction and a sliced contraction inside a recording"""
    rng = np.random.default_rng(11)
    ca = rng.uniform(-1, 1, (9, 14))
    x = rng.uniform(-1, 1, (9, 5))
    dca, dx = sip.DeviceBlock.from_numpy(ca), sip.DeviceBlock.from_numpy(x)
    with sip.recording():
        s = sip.slice_block(dca, (9, 5), (0, 4))
        s.axpy(dx, 2.0)
        sip.insert_block(dca, s, (0, 4))
        dca.scale(0.5)
        s2 = sip.slice_block(dca, (9, 5), (0, 4))
        got_s2 = s2.to_numpy()
    want = ca.copy()
    want[:, 4:9] += 2.0 * x
    want *= 0.5
    assert rel(dca.to_numpy(), want) <= 1e-15
    assert rel(got_s2, want[:, 4:9]) <= 1e-15


# ---------------------------------------------------------------------------------------------------------------
# the SIAL front-end on the device backend
# ---------------------------------------------------------------------------------------------------------------
def device_arrays(sip, host_arrays, kinds, segs):
    out = {}
    for name, blocks in host_arrays.items():
        A = sip.DistArray([segs[k] for k in kinds[name]])
        for idx, b in blocks.items():
            view = A.block_view(idx)
            sip._check(sip.lib().sipgpu_h2d(view.ptr, sip._hp(np.asfortranarray(b)), view.size), "h2d")
        sip.sync()
        out[name] = A
    return out


@pytest.mark.parametrize("segs", [{"o": [4, 4, 5], "v": [6, 7, 6]}, {"o": [20, 20], "v": [50, 50]}])
def test_lccd_sial_program_device_vs_oracle(sip, oracle, segs):
    from aces4_b200.sial_frontend import DeviceBackend, Program, Walker
    from sial_oracle_backend import OracleBackend
    from test_sial_frontend_cpu import KINDS, LCCD, make_arrays

    host = make_arrays(oracle, segs)
    small = sum(segs["v"]) < 40
    if small:
        ref_arrays = {k: {i: b.copy() for i, b in v.items()} for k, v in host.items()}
        be_o = OracleBackend(oracle, ref_arrays)
        e_ref = be_o.value(Walker(Program(LCCD), be_o, segs).run()["ecorrab"])
    results = {}
    for record in (False, True):
        arrays = device_arrays(sip, host, KINDS, segs)
        be = DeviceBackend(sip, arrays, record=record)
        l0 = sip.kernel_launches()
        scal = Walker(Program(LCCD), be, segs).run()
        e = be.value(scal["ecorrab"])
        launches = sip.kernel_launches() - l0
        blocks = {idx: arrays["t2new_ab"].get(idx).to_numpy() for idx in host["t2new_ab"]}
        results[record] = (blocks, e, launches, be.stats)
        for A in arrays.values():
            A.destroy()
    (b0, e0, l0, _), (b1, e1, l1, st) = results[False], results[True]
    for idx in b0:
        assert rel(b1[idx], b0[idx]) <= TOL
        if small:
            assert rel(b1[idx], ref_arrays["t2new_ab"][idx]) <= TOL
    assert abs(e1 - e0) <= TOL * abs(e0)
    if small:
        assert abs(e1 - e_ref) <= TOL * abs(e_ref)
    assert l1 * 4 < l0, (l0, l1)      # the recorded stream needs far fewer launches than op-at-a-time
    assert sum(s["chains"] for s in st) > 0 and sum(s["fused_accumulates"] for s in st) > 0


@pytest.mark.parametrize("dat", ["lccd_test.dat", "ccsdpt_test.dat", "eom_ccsd_water_test.dat", "lccd_frozencore_test.dat",
                                 "eom_water_dimer_test.dat", "second_ccsdpt_test.dat"])
def test_lccd_sial_program_at_the_shapes_of_the_shipped_inputs(sip, oracle, dat):
    """configs 1-3 of BASELINE.json: the energies need the integral/SCF stack (not buildable here), but the
    block SHAPES of those runs come from the .dat segment tables (tests/golden/dat_segments.json): the LCCD doubles
    program at exactly those occupied/virtual segments, seeded inputs, device (recorded) vs oracle at 1e-10."""
    import json

    from aces4_b200.sial_frontend import DeviceBackend, Program, Walker
    from sial_oracle_backend import OracleBackend
    from test_sial_frontend_cpu import KINDS, LCCD, make_arrays

    g = json.load(open(os.path.join(HERE, "golden", "dat_segments.json")))[dat]
    segs = {"o": g["occ"], "v": g["virt"]}
    host = make_arrays(oracle, segs)
    ref_arrays = {k: {i: b.copy() for i, b in v.items()} for k, v in host.items()}
    be_o = OracleBackend(oracle, ref_arrays)
    e_ref = be_o.value(Walker(Program(LCCD), be_o, segs).run()["ecorrab"])
    arrays = device_arrays(sip, host, KINDS, segs)
    be = DeviceBackend(sip, arrays, record=True)
    e = be.value(Walker(Program(LCCD), be, segs).run()["ecorrab"])
    for idx, want in ref_arrays["t2new_ab"].items():
        assert rel(arrays["t2new_ab"].get(idx).to_numpy(), want) <= TOL
    assert abs(e - e_ref) <= 1e-9 * max(1.0, abs(e_ref))      # north_star: 1e-9 on energies
    for A in arrays.values():
        A.destroy()


def test_sliced_contractions_inside_a_recording(sip, oracle):
    """half transformation with a static-array slice (T[a,i,mu,j] = T2[a,i,b,j]*ca[mu,b], read in place), accumulated
    into a slice of a larger array -- recorded: the sliced contractions are scheduled as opaque ops with the byte
    ranges of their parent arrays, so everything that touches the same parent stays in program order."""
    rng = np.random.default_rng(33)
    norb, nmo, v, o = 30, 40, 10, 6
    ca = np.asfortranarray(rng.uniform(-1, 1, (norb, nmo)))
    T2 = np.asfortranarray(rng.uniform(-1, 1, (v, o, v, o)))
    big = np.asfortranarray(rng.uniform(-1, 1, (v, o, norb, o)))
    dca, dT2, dbig = (sip.DeviceBlock.from_numpy(x) for x in (ca, T2, big))
    ptrn, _ = sip.get_contraction_ptrn([1, 2, 5, 4], [1, 2, 3, 4], [5, 3])
    want = big.copy(order="F")
    with sip.recording() as rec:
        for mu0, nmu, b0 in ((0, 10, 4), (10, 10, 14), (20, 10, 24)):
            sip.contract_sliced(ptrn, dT2, (v, o, v, o), None, dca, (nmu, v), (mu0, b0), (v, o, nmu, o), out=dbig,
                                dbeg=(0, 0, mu0, 0), alpha=0.5, beta=1.0)
            blk, ierr = oracle.block_slice(ca, (nmu, v), (mu0, b0))
            ref, ierr = oracle.block_contract(ptrn, T2, blk, (v, o, nmu, o))
            want[:, :, mu0:mu0 + nmu, :] += 0.5 * ref
        dbig.scale(2.0)
    assert rec.stats["recorded"] == 4 and rec.stats["levels"] == 4     # all four touch the same parent array: serial
    assert rel(dbig.to_numpy(), 2.0 * want) <= TOL


def test_repeated_stream_is_replayed_not_rescheduled(sip, oracle):
    """Every iteration of a CC program records the same pardo body on the same blocks.  A repeated recording of an
    identical stream must be served from the captured launches (sipgpu_wl_replays), give the same result on NEW operand
    values, and a stream that differs in one pointer must be scheduled afresh.  The temp addresses are a function of the
    pool's free lists: the FIRST recording may have to carve new arena space for some of its temps, so the second one (which
    finds all of them in the free list and takes them in address order) can still differ from it; from then on the stream
    is stable -- the second pass may or may not replay, the third and fourth must."""
    v, o, n = 10, 6, 3
    rng = np.random.default_rng(17)
    dlab, llab, rlab = [1, 2, 3, 4], [1, 5, 3, 6], [2, 5, 4, 6]
    dT2 = [[sip.DeviceBlock((v, o, v, o)) for _ in range(n)] for _ in range(n)]
    dV = [[sip.DeviceBlock((o, o, o, o)) for _ in range(n)] for _ in range(n)]
    D = sip.DeviceBlock((v, o, v, o))
    other = sip.DeviceBlock((v, o, v, o))

    def body(dest):
        dest.fill(0.0)
        for i1 in range(n):
            for j1 in range(n):
                t = sip.DeviceBlock((v, o, v, o))
                sip.contract_labels(dlab, (v, o, v, o), llab, dT2[i1][j1], rlab, dV[i1][j1], out=t)
                dest.accumulate(t)
                t.free()
        p = sip.DeviceBlock((v, o, v, o))
        sip.permute_labels([3, 2, 1, 4], [1, 2, 3, 4], dest, out=p)      # a permute and an elementwise op in the stream too
        dest.axpy(p, 0.5)
        p.free()

    replays = []
    last = 4
    for it in range(last + 1):
        T2 = [[rng.uniform(-1, 1, (v, o, v, o)) for _ in range(n)] for _ in range(n)]
        V = [[rng.uniform(-1, 1, (o, o, o, o)) for _ in range(n)] for _ in range(n)]
        want = np.zeros((v, o, v, o), order="F")
        for i1 in range(n):
            for j1 in range(n):
                sip._check(sip.lib().sipgpu_h2d(dT2[i1][j1].ptr, sip._hp(np.asfortranarray(T2[i1][j1])), dT2[i1][j1].size))
                sip._check(sip.lib().sipgpu_h2d(dV[i1][j1].ptr, sip._hp(np.asfortranarray(V[i1][j1])), dV[i1][j1].size))
                t, ierr = oracle.contract_labels(dlab, [v, o, v, o], llab, np.asfortranarray(T2[i1][j1]), rlab, np.asfortranarray(V[i1][j1]))
                assert ierr == 0
                want += t
        sip.sync()
        want = want + 0.5 * np.transpose(want, (2, 1, 0, 3))
        dest = other if it == last else D         # the last pass writes another destination: a different stream
        with sip.recording() as rec:
            body(dest)
            sip.wl_flush()
            replays.append(sip.wl_replays())
        assert rel(dest.to_numpy(), want) <= TOL, it
    assert replays[0] == 0 and replays[2] == 1 and replays[3] == 1 and replays[last] == 0, replays


@pytest.mark.parametrize("mode", [0, 3], ids=["ldg_red", "tma_bulk"])
def test_section_of_gets_and_accumulates_in_one_call(sip, mode):
    """sipgpu_array_get_many / _put_accumulate_many: a barrier section's worth of block traffic as one launch.  With
    copy_bulk = 3 whole blocks travel as TMA bulk transfers (cp.async.bulk into a shared-memory ring and out again;
    put += as cp.reduce.async.bulk add.f64), with 0 as LDG.128 / red.global.add.f64.  Run lengths straddle the 32 KB
    stages (one element short, odd counts and runs under the 64 KB minimum fall back to the register kernel)."""
    sizes = [8192, 8194, 8191, 12290, 4096, 20482, 40960, 3, 65536 + 2, 16384]
    a = sip.DistArray([sizes])
    rng = np.random.default_rng(5)
    want = [rng.standard_normal(n) for n in sizes]
    blocks = [(k + 1,) for k in range(len(sizes))]
    sip.set_tuning("copy_bulk", mode)
    try:
        for b, wv in zip(blocks, want):
            a.put(b, sip.DeviceBlock.from_numpy(wv))
        outs = [sip.DeviceBlock((n,)) for n in sizes]
        l0 = sip.kernel_launches()
        a.get_many(blocks, outs)
        assert 1 <= sip.kernel_launches() - l0 <= 2          # bulk-eligible runs + the rest
        for o, wv in zip(outs, want):
            assert np.array_equal(o.to_numpy().ravel(), wv)
        adds = [sip.DeviceBlock.from_numpy(np.full(n, 0.5 + k)) for k, n in enumerate(sizes)]
        a.put_accumulate_many(blocks, adds)
        a.put_accumulate_many(blocks, adds)
        for k, (b, wv) in enumerate(zip(blocks, want)):
            assert np.array_equal(a.get(b).to_numpy().ravel(), (wv + (0.5 + k)) + (0.5 + k))
    finally:
        sip.set_tuning("copy_bulk", -1)
        a.destroy()


def test_scale_after_its_producer_folds_into_alpha(sip, oracle):
    """scheduler pass A0 on numbers: `T = L*R; T *= -2; D1 += T; T2 = permute(T); D2 += T2` (the ph-ring body of rlccd_rhf.sialx)
    recorded == op-at-a-time == oracle"""
    rng = np.random.default_rng(11)
    v, o = 7, 5
    dl, ll, rl = [1, 2, 3, 4], [1, 5, 6, 4], [3, 6, 5, 2]
    ptrn, ierr = sip.get_contraction_ptrn(dl, ll, rl)
    assert ierr == 0
    L = np.asfortranarray(rng.uniform(-1, 1, (v, o, v, o)))
    R = np.asfortranarray(rng.uniform(-1, 1, (v, v, o, o)))
    ref, e = oracle.contract_labels(dl, [v, o, v, o], ll, L, rl, R)
    assert e == 0
    ref = -2.0 * ref.reshape((v, o, v, o), order="F")
    out = {}
    for mode in ("recorded", "eager"):
        dL, dR = sip.DeviceBlock.from_numpy(L), sip.DeviceBlock.from_numpy(R)
        D1, D2 = sip.DeviceBlock((v, o, v, o)).fill(1.0), sip.DeviceBlock((v, o, v, o)).fill(2.0)

        def body():
            T = sip.DeviceBlock((v, o, v, o))
            sip.contract(ptrn, dL, dR, (v, o, v, o), out=T)
            T.scale(-2.0)
            T2 = sip.DeviceBlock((v, o, v, o))
            sip.permute_labels([3, 4, 1, 2], [1, 2, 3, 4], T, out=T2)
            D1.accumulate(T)
            D2.accumulate(T2)
            T.free()
            T2.free()
        if mode == "recorded":
            with sip.recording() as rec:
                body()
            assert rec.stats["scheduled"] <= 3      # contraction (alpha = -2), D1 += T, fused permute-accumulate into D2
        else:
            body()
        out[mode] = (D1.to_numpy(), D2.to_numpy())
    for mode in out:
        assert np.max(np.abs(out[mode][0] - (1.0 + ref))) < 1e-12
        assert np.max(np.abs(out[mode][1] - (2.0 + ref.transpose(2, 3, 0, 1)))) < 1e-12
    assert np.array_equal(out["recorded"][0], out["eager"][0]) and np.array_equal(out["recorded"][1], out["eager"][1])
