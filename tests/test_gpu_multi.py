"""Two ranks on two GPUs of one box (skipped when fewer are visible): the IPC-mapped slabs of a distributed array --
get / put / put_accumulate / put_initialize across ranks over NVLink peer memory, the race detector at the barrier --
and the synthetic CCSD iteration at world 2 against world 1.  Rendezvous over gloo on 127.0.0.1."""
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

        torch.cuda.set_device(rank)
        sip.init(rank)
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
        A.destroy()
        # ---- the synthetic CCSD iteration, world 2 ----
        w = SyntheticCCSD([3, 3], [6, 6, 6], rank, world, exchange, dist.barrier, allreduce)
        e = w.iterate()
        chk = allreduce(w.t2new_checksum())
        q.put((rank, e, chk))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_on_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
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
