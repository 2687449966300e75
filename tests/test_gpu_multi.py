"""Two ranks = two processes, one per GPU when the box has two, BOTH ON GPU 0 otherwise (the CUDA-IPC slabs, the
cross-process `red.global.add.f64` accumulates, the peer-mapped gets and puts and the race detector at the barrier are
the same code either way; only the wire differs: NVLink peer memory vs. the same HBM through another process's mapping):
get / put / put_accumulate / put_initialize / put_increment across ranks, eagerly and inside a recording, and the
synthetic CCSD iteration at world 2 against world 1 with sampled blocks re-derived locally (the check bench.py prints as
`parity_vs_n1`).  Rendezvous over gloo on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aces4_b200 as sip
        from aces4_b200.sial_workload import SyntheticCCSD

        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        sip.init(dev)
        api = sip.api

        def exchange(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

        def allreduce(x):
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        # ---- block traffic between the ranks ----
        A = api.DistArray([[3, 4], [2, 5]], rank, world, exchange)
        A.track_accesses(True)
        blocks = [(i, j) for i in (1, 2) for j in (1, 2)]
        for idx in blocks:                               # every rank initialises the blocks it does NOT own (peer stores)
            if A.owner(idx) != rank:
                A.put_initialize(idx, 10.0 * A.block_number(idx))
        api.sync()
        dist.barrier()
        entries = exchange(A.section_accesses())
        api.consistency_validate([(b, f, r) for r, ent in enumerate(entries) for b, f in ent])   # disjoint writers: legal
        A.section_reset()
        for idx in blocks:                               # many-writer accumulate into every block from both ranks
            t = api.DeviceBlock(A.block_shape(idx)).fill(float(rank + 1))
            A.put_accumulate(idx, t)
        api.sync()
        dist.barrier()
        for idx in blocks:                               # every rank reads every block (local or peer)
            got = A.get(idx).to_numpy()
            assert np.all(got == 10.0 * A.block_number(idx) + 3.0), (rank, idx, got.ravel()[:3])
        dist.barrier()
        # many-writer section mixing put_increment and put += on the SAME blocks from both ranks (both count as
        # PUT_ACCUMULATE for the race rules, distributed_block_consistency.cpp:60; the reference serialises them at the
        # server, here both are red.global.add.f64): eagerly, then the same inside a recording.  No update may be lost.
        A.section_reset()
        expect = 10.0 * np.array([A.block_number(idx) for idx in blocks]) + 3.0
        for recorded in (False, True):
            if recorded:
                api.wl_begin()
            for rep in range(8):
                for idx in blocks:
                    A.put_increment(idx, 0.5 + rank)
                    t = api.DeviceBlock(A.block_shape(idx)).fill(0.25)
                    A.put_accumulate(idx, t)
                    if recorded:
                        t.free()
            if recorded:
                api.wl_end()
            api.sync()
            dist.barrier()
            entries = exchange(A.section_accesses())
            assert all(f == api.ACCESS_PUT_ACCUMULATE for ent in entries for _, f in ent)
            api.consistency_validate([(b, f, r) for r, ent in enumerate(entries) for b, f in ent])   # all-accumulate: legal
            A.section_reset()
            expect = expect + 8 * ((0.5 + 0) + (0.5 + 1) + 2 * 0.25)
            for k, idx in enumerate(blocks):
                got = A.get(idx).to_numpy()
                assert np.all(got == expect[k]), (rank, recorded, idx, got.ravel()[:3], expect[k])
            dist.barrier()
            A.section_reset()
        # one section's worth of put += and of gets in one call each (sipgpu_array_put_accumulate_many / _get_many: one launch)
        srcs = [api.DeviceBlock(A.block_shape(idx)).fill(1.0 + k) for k, idx in enumerate(blocks)]
        A.put_accumulate_many(blocks, srcs)
        api.sync()
        dist.barrier()
        outs = [api.DeviceBlock(A.block_shape(idx)) for idx in blocks]
        A.get_many(blocks, outs)
        api.sync()
        for k, o in enumerate(outs):
            assert np.all(o.to_numpy() == expect[k] + 2 * (1.0 + k)), (rank, k)
        dist.barrier()
        A.section_reset()
        # create -> first remote put with no barrier in between: the owner's zero fill must already have landed
        B = api.DistArray([[64, 64], [64]], rank, world, exchange)
        for idx in ((1, 1), (2, 1)):
            if B.owner(idx) != rank:
                B.put_accumulate(idx, api.DeviceBlock(B.block_shape(idx)).fill(7.0))
        api.sync()
        dist.barrier()
        for idx in ((1, 1), (2, 1)):
            assert np.all(B.get(idx).to_numpy() == 7.0)
        dist.barrier()
        B.destroy()
        A.destroy()
        # ---- the synthetic CCSD iteration, world 2 ----
        w = SyntheticCCSD([3, 3], [6, 6, 6], rank, world, exchange, dist.barrier, allreduce)
        e = w.iterate()
        chk = allreduce(w.t2new_checksum())
        worst = w.verify_blocks(12, offset=rank * 5)   # owner's block (peer get) vs. recomputed from replicated inputs
        assert worst <= 1e-12, worst
        q.put((rank, e, chk))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_share_distributed_arrays():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    import aces4_b200 as sip
    from aces4_b200.sial_workload import SyntheticCCSD

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    sip.init(0)
    w1 = SyntheticCCSD([3, 3], [6, 6, 6])
    e1 = w1.iterate()
    chk1 = w1.t2new_checksum()
    for _, e, chk in got:
        assert abs(e - e1) <= 1e-9 * max(1.0, abs(e1))
        assert abs(chk - chk1) <= 1e-10 * chk1
