"""CPU tests of the deferred op stream's scheduler (aces4_b200/csrc/worklist.cu) in DRY mode: ops are recorded with
fake device addresses and scheduled on the host only.  What is checked is the host logic: temp forwarding
(T = L*R; D += T  ->  D += L*R), chain fusion, zero-fill elimination, and -- on random op streams -- that the levels
respect every read/write hazard of the recorded program order (an independent O(n^2) checker written here).
No GPU needed; the numerical equivalence of recorded vs op-at-a-time execution is tests/test_gpu_worklist.py."""
import random

import pytest


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


class Rec:
    """Mirror of the recorded stream kept by the test: (reads, writes) address ranges per op."""

    def __init__(self, sip):
        self.sip, self.ops = sip, []

    def blk(self, shape):
        return self.sip.DeviceBlock(shape)

    @staticmethod
    def rng(b):
        return (b.ptr, b.ptr + 8 * b.size)

    def note(self, reads, writes):
        self.ops.append(([self.rng(b) for b in reads], [self.rng(b) for b in writes]))

    def fill(self, d, v):
        d.fill(v)
        self.note([], [d])

    def scale(self, d, f):
        d.scale(f)
        self.note([d], [d])

    def axpy(self, d, s, f):
        d.axpy(s, f)
        self.note([d, s], [d])

    def copy(self, d, s, f=1.0):
        d.scale_and_copy(s, f)
        self.note([s], [d])

    def add_sub(self, d, l, r, sign):
        d.set_add_sub(l, r, sign)
        self.note([l, r], [d])

    def permute(self, d, s, transp):
        self.sip.permute(s, transp, out=d)
        self.note([s], [d])

    def contract(self, ptrn, L, R, D, beta=0.0, alpha=1.0):
        self.sip.contract(ptrn, L, R, D.shape, out=D, alpha=alpha, beta=beta)
        self.note([L, R] + ([D] if beta != 0.0 else []), [D])


def overlap(a, b):
    return a[0] < b[1] and b[0] < a[1]


def check_plan(ops, level, unit):
    """every conflicting pair (RAW, WAR, WAW) of the recorded order that was not fused into one unit is level-ordered"""
    n = len(ops)
    assert len(level) == n and len(unit) == n
    for j in range(n):
        rj, wj = ops[j]
        for i in range(j):
            if unit[i] == unit[j]:
                continue
            ri, wi = ops[i]
            conflict = any(overlap(x, y) for x in wi for y in rj + wj) or any(overlap(x, y) for x in ri for y in wj)
            if conflict:
                assert level[i] < level[j], (i, j, level[i], level[j])


def test_hhladder_body_becomes_one_chain(sip):
    """rlccd_rhf.sialx:342-355: do i1, j1: T = T2old[a,i1,b,j1]*V[i,i1,j,j1]; Taibj += T  -> one chained problem"""
    v, o, nseg = 8, 5, 3
    ptrn, ierr = sip.get_contraction_ptrn([1, 2, 3, 4], [1, 5, 3, 6], [2, 5, 4, 6])
    assert ierr == 0
    with sip.recording(dry=True) as rec:
        T2 = [[sip.DeviceBlock((v, o, v, o)) for _ in range(nseg)] for _ in range(nseg)]
        V = [[sip.DeviceBlock((o, o, o, o)) for _ in range(nseg)] for _ in range(nseg)]
        D = sip.DeviceBlock((v, o, v, o))
        D.fill(0.0)
        for i1 in range(nseg):
            for j1 in range(nseg):
                T = sip.DeviceBlock((v, o, v, o))
                sip.contract(ptrn, T2[i1][j1], V[i1][j1], (v, o, v, o), out=T)
                D.accumulate(T)
                T.free()
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    assert st["recorded"] == 1 + 2 * nseg * nseg
    assert st["fused_accumulates"] == nseg * nseg and st["temps_elided"] == nseg * nseg
    assert st["chains"] == 1 and st["chain_pairs"] == nseg * nseg
    assert st["scheduled"] == 1 and st["levels"] == 1      # the zero fill disappeared into beta = 0
    assert len(set(unit)) == 1 and set(level) == {1}


def test_temp_with_second_reader_is_not_forwarded(sip):
    v = 6
    ptrn, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    with sip.recording(dry=True):
        A, B, D, E = (sip.DeviceBlock((v, v)) for _ in range(4))
        T = sip.DeviceBlock((v, v))
        sip.contract(ptrn, A, B, (v, v), out=T)
        D.accumulate(T)
        E.accumulate(T)       # a second consumer: T must be materialised
        T.free()
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    assert st["fused_accumulates"] == 0 and st["scheduled"] == 3
    assert level == [1, 2, 2]


def test_operand_overwritten_between_producer_and_consumer_blocks_fusion(sip):
    v = 6
    ptrn, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    with sip.recording(dry=True):
        A, B, D = (sip.DeviceBlock((v, v)) for _ in range(3))
        T = sip.DeviceBlock((v, v))
        sip.contract(ptrn, A, B, (v, v), out=T)
        A.fill(1.0)           # WAR on A: T must be computed from the OLD A
        D.accumulate(T)
        T.free()
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    assert st["fused_accumulates"] == 0
    assert level == [1, 2, 2]


def test_unfreed_block_is_not_elided(sip):
    v = 6
    ptrn, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    with sip.recording(dry=True):
        A, B, D, T = (sip.DeviceBlock((v, v)) for _ in range(4))
        sip.contract(ptrn, A, B, (v, v), out=T)
        D.accumulate(T)       # T stays live after the recording: its value is observable
        sip.wl_flush()
        st = sip.wl_stats()
    assert st["fused_accumulates"] == 0 and st["scheduled"] == 2


def test_permute_accumulate_fusion_and_batching(sip):
    """handle_block_add with differing labels: permute into a temp, then add (interpreter.cpp:1874-1997)"""
    shape = (4, 3, 4, 3)
    with sip.recording(dry=True):
        X = [sip.DeviceBlock(shape) for _ in range(5)]
        D = [sip.DeviceBlock(shape) for _ in range(5)]
        for x, d in zip(X, D):
            t = sip.DeviceBlock(shape)
            sip.permute(x, [1, 3, 2, 1, 4], out=t)
            d.axpy(t, -1.0)
            t.free()
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    assert st["fused_accumulates"] == 5 and st["scheduled"] == 5 and st["levels"] == 1
    assert set(level) == {1}


@pytest.mark.parametrize("head_assigns", [False, True])
def test_chain_skips_commuting_accumulates_of_other_extents(sip, head_assigns):
    """non-uniform contracted segments: D += A1*B1 (k=3); D += A2*B2 (k=5); D += A3*B3 (k=3).  The k=3 pair is one
    chain, the k=5 accumulate commutes with it -- unless the head ASSIGNS, which must not be moved past an accumulate."""
    v = 6
    ptrn, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    with sip.recording(dry=True):
        D = sip.DeviceBlock((v, v))
        ops = [(sip.DeviceBlock((v, k)), sip.DeviceBlock((k, v))) for k in (3, 5, 3)]
        for n, (a, b) in enumerate(ops):
            sip.contract(ptrn, a, b, (v, v), out=D, beta=0.0 if (n == 0 and head_assigns) else 1.0)
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    if head_assigns:
        assert st["chains"] == 0 and level == [1, 2, 3]
    else:
        assert st["chains"] == 1 and st["chain_pairs"] == 2 and unit == [2, 1, 2] and level == [2, 1, 2]


def test_pattern_error_surfaces_at_the_recording_call(sip):
    with sip.recording(dry=True):
        A, B, D = sip.DeviceBlock((4, 5)), sip.DeviceBlock((6, 4)), sip.DeviceBlock((4, 4))
        ptrn, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
        with pytest.raises(sip.SipGpuError):
            sip.contract(ptrn, A, B, (4, 4), out=D)   # contracted extents 5 vs 6
        assert sip.wl_stats()["recorded"] == 0


def test_nested_begin_is_an_error_and_compute_needs_a_device(sip):
    sip.wl_begin(dry=True)
    try:
        assert sip.lib().sipgpu_wl_recording() == 2
        assert sip.lib().sipgpu_wl_begin(1) == 105   # SIPGPU_E_STATE
    finally:
        sip.wl_end()
    assert sip.lib().sipgpu_wl_recording() == 0
    assert sip.lib().sipgpu_wl_end() == 105


def test_auto_flush_limit(sip):
    with sip.recording(dry=True):
        sip.lib().sipgpu_wl_set_limits(10, 0)
        a = sip.DeviceBlock((16,))
        for k in range(35):
            a.scale(1.5)
        st = sip.wl_stats()
    assert st["flushes"] >= 3 and st["recorded"] == 35


@pytest.mark.parametrize("seed", range(40))
def test_random_streams_respect_all_hazards(sip, seed):
    rnd = random.Random(seed)
    v = 4
    p2, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    with sip.recording(dry=True):
        r = Rec(sip)
        live = [r.blk((v, v)) for _ in range(6)]
        temps = []
        for step in range(160):
            k = rnd.randrange(9)
            pick = lambda: rnd.choice(live + temps)  # noqa: E731
            if k == 0:
                r.fill(pick(), 0.0 if rnd.random() < 0.5 else 2.0)
            elif k == 1:
                r.scale(pick(), 0.5)
            elif k == 2:
                d, s = pick(), pick()
                if d is not s:
                    r.axpy(d, s, rnd.choice([1.0, -1.0, 0.5]))
            elif k == 3:
                d, s = pick(), pick()
                if d is not s:
                    r.copy(d, s)
            elif k == 4:
                d, a, b = pick(), pick(), pick()
                r.add_sub(d, a, b, rnd.choice([1.0, -1.0]))
            elif k == 5:
                d, s = pick(), pick()
                if d is not s:
                    r.permute(d, s, [1, 2, 1])
            elif k in (6, 7):
                d, a, b = pick(), pick(), pick()
                if d is not a and d is not b:
                    r.contract(p2, a, b, d, beta=rnd.choice([0.0, 1.0, 1.0]))
            else:
                # the interpreter's temp idiom: T = A*B ; D += T ; free T
                a, b, d = pick(), pick(), rnd.choice(live)
                t = r.blk((v, v))
                r.contract(p2, a, b, t)
                r.axpy(d, t, 1.0)
                t.free()
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    assert st["recorded"] == len(r.ops)
    check_plan(r.ops, level, unit)
    assert st["scheduled"] <= st["recorded"] and st["levels"] >= 1


def test_lccd_pardo_stream_dry_recording_counts_and_host_cost(sip, tmp_path):
    """scripts/micro/wl_dry_bench.cpp: the op-at-a-time stream of two LCCD pardo bodies (34 992 ops at 3 x 6 segments)
    recorded in dry mode from C++.  The fusion counts are exact properties of that stream: every `T = L*R; D += T` pair of
    hhladder and both permuted accumulates' producers of phladder are forwarded (8 748 = 3 888 + 4 860 temps elided) and
    hhladder's 324 destinations each become one chain.  The host cost per recorded op is asserted loosely (it is a
    few tenths of a microsecond; the bound only catches an accidental return to per-op heap traffic or quadratic passes)."""
    import json
    import os
    import subprocess

    import aces4_b200

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.dirname(aces4_b200.lib_path())
    exe = str(tmp_path / "wl_dry_bench")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "scripts", "micro", "wl_dry_bench.cpp"),
                           "-I", os.path.join(root, "include"), "-L", lib_dir, "-lsipgpu", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    out = json.loads(subprocess.run([exe, "16", "3", "16", "6", "3"], capture_output=True, text=True, check=True).stdout)
    assert out["ops_recorded"] == 34992
    assert out["fused_accumulates"] == 8748 and out["chains"] == 324
    assert out["ops_scheduled"] < out["ops_recorded"]
    assert out["us_per_op"] < 5.0, out


def test_scale_directly_after_its_producer_folds_into_alpha(sip):
    """pass A0: `T = L*R; T *= -1; D += T` (phladder_ab of rlccd_rhf.sialx) -- the whole-block scale is absorbed by the contraction,
    after which T has a single consumer and the accumulate fuses too: ONE scheduled op.  A scale that does not directly follow the
    producer (another reader in between) stays."""
    v = 6
    ptrn, _ = sip.get_contraction_ptrn([1, 2], [1, 3], [3, 2])
    with sip.recording(dry=True):
        A, B, D = (sip.DeviceBlock((v, v)) for _ in range(3))
        T = sip.DeviceBlock((v, v))
        sip.contract(ptrn, A, B, (v, v), out=T)
        T.scale(-1.0)
        D.accumulate(T)
        T.free()
        sip.wl_flush()
        level, unit = sip.wl_last_plan()
        st = sip.wl_stats()
    assert st["scheduled"] == 1 and st["fused_accumulates"] == 1 and unit[1] == unit[0] == 2, (st, unit)
    with sip.recording(dry=True):
        A, B, D, E = (sip.DeviceBlock((v, v)) for _ in range(4))
        T = sip.DeviceBlock((v, v))
        sip.contract(ptrn, A, B, (v, v), out=T)
        E.accumulate(T)       # reads the unscaled T first
        T.scale(-1.0)
        D.accumulate(T)
        T.free()
        sip.wl_flush()
        st = sip.wl_stats()
    assert st["scheduled"] == 4 and st["fused_accumulates"] == 0, st
