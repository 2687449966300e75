"""The reference's LCCD / CCSD energy goldens on N GPUs of one box: every rank walks the whole program and executes its
share of the pardo iterations (loop_manager.cpp:468-499); the served / distributed arrays are block-cyclic over the ranks
(CUDA-IPC slabs), remote blocks are fetched / accumulated over NVLink.  NOT yet run (written when the round's GPU minutes
were spent); the same worker partition on shared arrays is checked on the CPU by
tests/test_lccd_water_energy_cpu.py::test_programs_on_several_workers_reproduce_the_goldens.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_n_energy.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lccd_water as lw  # noqa: E402
from aces4_b200.sial_frontend import DeviceBackend, Program, Walker  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import aces4_b200
    torch.cuda.set_device(local)
    aces4_b200.init(local)
    sip = aces4_b200.api

    def exchange(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def allreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()

    ok = True
    for program, text, case, golden in (("lccd", lw.PROGRAM, "fine", lw.GOLDEN["lccd_energy"]),
                                        ("ccsd", lw.PROGRAM_CCSD, "all_fine", lw.golden_ccsd()[0])):
        inp = lw.inputs(case)
        sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
        arrays = {}
        for name, kinds in lw.KINDS.items():
            A = sip.DistArray([inp["segs"][k] for k in kinds], rank, world, exchange if world > 1 else None)
            A.fill_local(0.0)
            for idx, b in inp["arrays"][name].items():
                if A.owner(idx) == rank:            # every rank uploads the blocks it owns
                    v = A.block_view(idx)
                    sip._check(sip.lib().sipgpu_h2d(v.ptr, sip._hp(np.asfortranarray(b)), v.size), "h2d")
            arrays[name] = A
        sip.sync()
        barrier()
        be = DeviceBackend(sip, arrays, record=True, rank=rank, world=world, barrier=barrier if world > 1 else None,
                           allreduce=allreduce if world > 1 else None)
        be.fock = sip.DeviceBlock.from_numpy(inp["fock"])
        w = Walker(Program(text), be, inp["segs"], rank=rank, world=world, index_base=inp["index_base"])
        t0 = time.perf_counter()
        _, hist = lw.converge(w, be.value, max_iter=150)
        dt = time.perf_counter() - t0
        e = hist[-1] + inp["e_scf"]
        good = abs(e - golden) < lw.GOLDEN["tolerance"]
        ok = ok and good
        if rank == 0:
            print(json.dumps({"program": program, "case": case, "n_gpus": world, "energy": e, "golden": golden,
                              "diff": e - golden, "ok": good, "iterations": len(hist), "seconds": round(dt, 2)}), flush=True)
        barrier()
        for A in arrays.values():
            A.destroy()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
