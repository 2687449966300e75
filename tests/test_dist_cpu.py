"""N > 1 host-side logic on CPU: block ownership / partition of destinations (against the oracle's restatement of
array_table.cpp + data_distribution.cpp) and the rendezvous plumbing bench.py uses (handle exchange, collective_sum,
max-over-ranks timing) under a world_size-2 gloo group.  The device half (IPC slabs, peer copies) is covered by the
-m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_matches_oracle_ownership(oracle):
    import aces4_b200 as sip

    sip.build()
    api = sip.api
    for nseg in ([12, 3, 12, 3], [2, 3], [3, 3, 3, 3], [5], [2, 2, 2, 2, 2, 2]):
        for world in (1, 2, 4, 8):
            parts = api.partition_blocks(nseg, world)
            seen = set()
            for r, blocks in enumerate(parts):
                for idx in blocks:
                    num = oracle.block_number(nseg, [1] * len(nseg), list(idx))
                    assert api.layout_block_number(nseg, idx) == num
                    assert oracle.block_owner(num, world) == r == api.layout_block_owner(num, world)
                    assert oracle.block_num2id(nseg, [1] * len(nseg), num) == list(idx)  # check_block_number_calc
                    seen.add(idx)
            assert len(seen) == int(np.prod(nseg))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1  # block-cyclic balance
    assert api.layout_block_number([2, 3], [3, 1]) == -1  # out of range segment


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aces4_b200 as sip

        api = sip.api
        # handle exchange as DistArray does it: every rank contributes 64 opaque bytes, gets all of them in rank order
        mine = bytes([rank]) * 64
        out = [None] * world
        dist.all_gather_object(out, mine)
        assert out == [bytes([r]) * 64 for r in range(world)]
        # destinations this rank computes + collective_sum of a per-rank partial (sial_ops_parallel.cpp:549-565)
        nseg = [4, 2, 4, 2]
        blocks = api.partition_blocks(nseg, world)[rank]
        partial = float(sum(api.layout_block_number(nseg, b) for b in blocks))
        t = torch.tensor([partial], dtype=torch.float64)
        dist.all_reduce(t)
        nb = int(np.prod(nseg))
        assert t.item() == nb * (nb - 1) / 2
        # max-over-ranks timing reduction used by bench.py
        m = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        assert m.item() == 10.0 + world - 1
        # sip_barrier with race detection: every rank contributes what it touched in the section; all ranks validate the
        # union (distributed_block_consistency.cpp rules).  Section 1: put += from both ranks into the same blocks and
        # disjoint puts -> legal.  Section 2: rank 0 puts block 5, rank 1 gets it without a barrier -> inconsistent.
        def barrier_validate(mine):
            out = [None] * world
            dist.all_gather_object(out, mine)
            api.consistency_validate([(b, bits, r) for r, ent in enumerate(out) for b, bits in ent])

        barrier_validate([(b, api.ACCESS_PUT_ACCUMULATE) for b in range(8)] + [(100 + rank, api.ACCESS_PUT)])
        try:
            barrier_validate([(5, api.ACCESS_PUT if rank == 0 else api.ACCESS_GET)])
            raced = False
        except api.SipGpuError:
            raced = True
        assert raced
        dist.barrier()
        q.put((rank, len(blocks)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_plumbing():
    import aces4_b200 as sip

    sip.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert got == {0: 32, 1: 32}
