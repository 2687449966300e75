"""Block traffic between two ranks (two processes; one GPU each when the box has two, else both on GPU 0): GB/s per GPU of a
section of `get`s (sipgpu_array_get_many) and of `put +=` (sipgpu_array_put_accumulate_many) towards the neighbour's slab,
with whole-block copies as TMA bulk transfers (copy_bulk 3) and as LDG.128 / red.global.add.f64 (copy_bulk 0); both ranks at
once and one rank alone; exactness of what arrives is checked.  Writes gpurun_out/peer_traffic.jsonl."""
import json
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aces4_b200 as sip

    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    sip.init(dev)
    api = sip.api
    stream = torch.cuda.ExternalStream(api.stream_handle())

    def exchange(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows = []
    cases = ((100, 10), (50, 24), (16, 64))                   # 800 MB, 50 MB and 512 KB blocks
    if os.environ.get("PROBE_QUICK"):
        cases = ((50, 24),)
    for seg, nseg in cases:
        shape = [[seg] * nseg, [seg] * 2, [seg], [seg]] if seg != 16 else [[16] * nseg, [16] * 8, [16], [16]]
        A = api.DistArray(shape, rank, world, exchange)
        blocks = [(i, j, 1, 1) for i in range(1, len(shape[0]) + 1) for j in range(1, len(shape[1]) + 1)]
        peer = [b for b in blocks if A.owner(b) == (rank + 1) % world] if world > 1 else blocks
        tmp = [api.DeviceBlock(A.block_shape(b)) for b in peer]
        nbytes = 8.0 * sum(int(np.prod(A.block_shape(b))) for b in peer)
        sec = A.section(peer, tmp)
        for mode in (0, 3):
            api.set_tuning("copy_bulk", mode)
            for who in ("all", "rank0"):
                for op in ("get", "put_acc"):
                    # contents: owner fills block k with k + 1; put += adds 0.5 (exact)
                    for b in blocks:
                        if A.owner(b) == rank:
                            A.block_view(b).fill(float(A.block_number(b) + 1))
                    for t in tmp:
                        t.fill(0.5)
                    api.sync()
                    dist.barrier()
                    active = who == "all" or rank == 0

                    def go():
                        if not active:
                            return
                        if op == "get":
                            A.get_many(section=sec)
                        else:
                            A.put_accumulate_many(section=sec)
                    ts = []
                    for rep in range(4):
                        api.sync()
                        dist.barrier()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        go()
                        e1.record(stream)
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                        dist.barrier()
                    ms = reduce_max(float(np.median(ts[1:]))) if who == "all" else float(np.median(ts[1:]))
                    api.sync()
                    dist.barrier()
                    ok = True
                    if op == "get" and active:
                        for b, t in list(zip(peer, tmp))[:: max(1, len(peer) // 7)]:
                            ok &= bool(np.all(t.to_numpy() == A.block_number(b) + 1))
                    if op == "put_acc":
                        for b in blocks[:: max(1, len(blocks) // 9)]:
                            written = 2.0 if (world == 1 or who == "all" or A.owner(b) == 1) else 0.0
                            got = A.get(b).to_numpy()
                            ok &= bool(np.all(got == A.block_number(b) + 1 + written))
                    if rank == 0:
                        rows.append({"block_bytes": 8 * int(np.prod(A.block_shape(peer[0]))), "blocks": len(peer), "op": op,
                                     "copy_bulk": mode, "ranks_active": who, "ms": ms, "GBps_per_gpu": nbytes / ms / 1e6,
                                     "exact": ok, "world": world, "gpus": ngpu})
                        print(rows[-1], flush=True)
        api.sync()
        dist.barrier()
        del tmp
        A.destroy()
    api.set_tuning("copy_bulk", -1)
    if rank == 0:
        q.put(rows)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp

    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows = q.get(timeout=600)
    for p in procs:
        p.join(60)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"peer_traffic_w{world}.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    assert all(r["exact"] for r in rows), [r for r in rows if not r["exact"]]
