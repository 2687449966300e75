#!/usr/bin/env python
"""bench.py -- FP64 block-contraction throughput and CCSD iteration time (BASELINE.json metric) on N B200s.

A "step" is one synthetic CCSD (LCCD-shaped) iteration over block-sparse distributed arrays at o = 60 (3 x 20),
v = 600 (12 x 50) -- config 5 of BASELINE.json, 1.2226 PFLOP of algorithmic contraction work -- executed by
aces4_b200/sial_workload.py through the C ABI of libsipgpu.so.  Strong scaling: the problem is fixed, destination
blocks are partitioned block-cyclically over the ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                           (the reference's CPU algorithm on the host cores)

Prints ONE JSON line (rank 0).  `value` = algorithmic TFLOP/s with inputs resident in HBM; `e2e` = the same with the
iteration's input (this rank's T2old slab) coming from pinned host memory and its result (T2new slab + energy)
going back every step; `roofline` = the contraction kernel's achieved FP64 TFLOP/s (CUDA events around its launches
inside the timed steps) against the DMMA issue-rate peak measured in the same run; `cpu_baseline` = the oracle
(CPU restatement of tensor_block_contract_, permute -> OpenBLAS dgemm -> permute, all host cores) on a bounded sample
of the same work-lists.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# The CPU arm runs two thread pools in turn -- the oracle's OpenMP loops (permutes, as in the reference's Fortran) and
# OpenBLAS's own threads (dgemm).  With libgomp's default active wait policy the idle OpenMP workers spin through the
# first milliseconds of every dgemm and take the cores away from it (measured: a pp-ladder block pair 82 ms vs 35 ms).
# A reference build has one OpenMP runtime for both and no such conflict, so the baseline must not have it either.
# Must be set before libgomp is loaded (torch, liboracle.so), hence here.
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("GOMP_SPINCOUNT", "0")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

O_SEGS = [20, 20, 20]   # o = 60
V_SEGS = [50] * 12      # v = 600
METRIC = "FP64 block-contraction TFLOP/s (synthetic CCSD iteration; ms_per_step = CCSD iteration time)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sipgpu", choices=["sipgpu", "reference"])
    ap.add_argument("--o-segs", type=str, default=None, help="development only: e.g. 20,20 (default 20,20,20)")
    ap.add_argument("--v-segs", type=str, default=None, help="development only: e.g. 50,50,50,50")
    ap.add_argument("--cpu-dests", type=int, default=2, help="destination blocks per CPU sample")
    ap.add_argument("--density", type=float, default=1.0,
                    help="block density of the amplitude array (SURVEY 8d item 3 asks for 1.0 and 0.5); the headline is 1.0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(o_segs, v_segs, density=1.0):
    return (f"synthetic CCSD (LCCD-shaped, rlccd_rhf.sialx) iteration o={sum(o_segs)} ({len(o_segs)}x{o_segs[0]}) "
            f"v={sum(v_segs)} ({len(v_segs)}x{v_segs[0]}), block density {density:g}: hh ladder + 3 ph ring terms + pp(AO) ladder "
            f"stand-in + T2new symmetrisation (put_accumulate) + energy")


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure) timed as the reference's CPU implementation of the path
# ----------------------------------------------------------------------------------------------------------
class CpuSample:
    """A bounded sample of the iteration's work-lists on the host: `ndest` destination blocks, every term, chained over
    the contracted segments exactly like the SIAL body (contract into a temp block, accumulate)."""

    def __init__(self, o_segs, v_segs, ndest, threads):
        from oracle import oracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from workload_ref import RefWorkload
        from aces4_b200.sial_workload import TERMS

        self.oracle, self.TERMS = oracle, TERMS
        os.environ["OMP_NUM_THREADS"] = str(threads)
        self.blas = oracle.use_openblas(threads)
        self.o_segs, self.v_segs = o_segs, v_segs
        # block generator only (no dense assembly at full size)
        self.ref = RefWorkload.__new__(RefWorkload)
        self.ref.oracle, self.ref.o_segs, self.ref.v_segs = oracle, list(o_segs), list(v_segs)
        self.ref.seed, self.ref.ao_pool = 0xACE54, 8
        self.ref.segs = {"v": list(v_segs), "o": list(o_segs)}
        self.ref.kinds = {"T2old": "vovo", "Vvovo": "vovo", "Voooo": "oooo", "TY": "vovo", "Vovvo": "ovvo", "Vvvoo": "vvoo"}
        nv, no = len(v_segs), len(o_segs)
        all_blocks = [(a, i, b, j) for a in range(1, nv + 1) for i in range(1, no + 1) for b in range(1, nv + 1)
                      for j in range(1, no + 1)]
        stride = max(1, len(all_blocks) // ndest)
        self.dests = all_blocks[::stride][:ndest]
        self.cache = {}
        self.jobs = []   # (term, dlabels, dext, [(llabels, L, rlabels, R)])
        self.flops = 0.0
        isv = lambda c: c in "abcd"
        import numpy as np
        for blk in self.dests:
            for t in TERMS:
                dlab, llab, rlab = t["dlab"], t["llab"], t["rlab"]
                labs = sorted(set(dlab + llab + rlab))
                num = {c: n + 1 for n, c in enumerate(labs)}
                contracted = [c for c in llab if c in rlab]
                segs = dict(zip(dlab, blk))
                dext = [(v_segs if isv(c) else o_segs)[segs[c] - 1] for c in dlab]
                pairs = []
                for cseg in np.ndindex(*[len(v_segs if isv(c) else o_segs) for c in contracted]):
                    for c, s in zip(contracted, cseg):
                        segs[c] = s + 1
                    ops = []
                    for name, lab in ((t["L"], llab), (t["R"], rlab)):
                        ops.append(self._block(name, tuple(segs[c] for c in lab)))
                    pairs.append((ops[0], ops[1]))
                    self.flops += 2.0 * float(np.prod(dext)) * float(np.prod([ops[0].shape[llab.index(c)] for c in contracted]))
                self.jobs.append((t, [num[c] for c in dlab], dext, [num[c] for c in llab], [num[c] for c in rlab], pairs))

    def _block(self, name, idx):
        key = (name, idx)
        if name == "aoint":
            ext = tuple(self.v_segs[i - 1] for i in idx)
            key = ("aoint", ext, (idx[0] * 7 + idx[1] * 3 + idx[2] * 5 + idx[3]) % 8)
        if key not in self.cache:
            if name == "aoint":
                self.cache[key] = self.ref.ao_block(*idx)
            elif name == "W":
                c_, k_, a_, i_ = idx
                import numpy as np
                self.cache[key] = np.asfortranarray(self.ref.block("T2old", idx) -
                                                    np.transpose(self.ref.block("T2old", (c_, i_, a_, k_)), (0, 3, 2, 1)))
            else:
                self.cache[key] = self.ref.block(name, idx)
        return self.cache[key]

    def run_once(self):
        import numpy as np
        t0 = time.perf_counter()
        chk = 0.0
        for t, dl, dext, ll, rl, pairs in self.jobs:
            acc = np.zeros(dext, order="F")
            for L, R in pairs:
                d, ierr = self.oracle.contract_labels(dl, dext, ll, L, rl, R)
                assert ierr == 0
                acc += d
            chk += float(acc.ravel()[0])
        return time.perf_counter() - t0, chk


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, o_segs, v_segs):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    s = CpuSample(o_segs, v_segs, args.cpu_dests, cores)
    for _ in range(args.warmup):
        s.run_once()
    times = [s.run_once()[0] for _ in range(args.steps)]
    sec = sum(times) / len(times)
    tf = s.flops / sec / 1e12
    sample = (f"{len(s.dests)} of {len(v_segs) ** 2 * len(o_segs) ** 2} destination blocks x all 5 terms "
              f"({s.flops / 1e9:.0f} GFLOP per step), oracle permute->dgemm->permute + accumulate, {s.blas.split()[0]} dgemm")
    line = {"metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(o_segs, v_segs), "sample": sample},
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def main():
    args = parse_args()
    o_segs = [int(x) for x in args.o_segs.split(",")] if args.o_segs else O_SEGS
    v_segs = [int(x) for x in args.v_segs.split(",")] if args.v_segs else V_SEGS
    if args.impl == "reference":
        run_reference(args, o_segs, v_segs)
        return

    import numpy as np
    import torch
    import aces4_b200 as sip
    from aces4_b200.sial_workload import SyntheticCCSD, iteration_flops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run); got {world}")
    torch.cuda.set_device(local)
    sip.init(local)
    exchange = barrier = allreduce = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def exchange(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

        def barrier():
            dist.barrier()

        def allreduce(x):
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            return float(t.item())

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        return x if world == 1 else allreduce(x)

    api = sip.api
    stream = torch.cuda.ExternalStream(api.stream_handle())
    peak_tf = max(api.dmma_peak_probe(40000) for _ in range(2))  # FP64 tensor (DMMA) issue-rate peak, this GPU, this run

    w = SyntheticCCSD(o_segs, v_segs, rank, world, exchange, barrier or (lambda: None), allreduce, density=args.density)
    # algorithmic flops of one iteration: the operand pairs that exist (all of them at density 1.0)
    flops = iteration_flops(o_segs, v_segs) if args.density >= 1.0 else reduce_sum(w.flops)

    def fence():
        sip.sync()
        if world > 1:
            barrier()
        torch.cuda.synchronize()

    def timed(nsteps, body):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = None
        for _ in range(nsteps):
            res = body()
        e1.record(stream)
        e1.synchronize()
        fence()
        return reduce_max(e0.elapsed_time(e1)), res

    for _ in range(args.warmup):
        energy = w.iterate()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local) if rank == 0 else None
    from aces4_b200.sial_workload import term_flops
    kev, pending = [], []   # (flops of this rank, e0, e1) around every launch of the contraction kernel

    def before_launch(t):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        pending.append(e)

    def after_launch(t):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        kev.append((w.term_flops_rank.get(t["name"], 0.0), pending.pop(), e))

    w.before_launch, w.after_launch = before_launch, after_launch
    launches0 = sip.kernel_launches()
    ms_total, energy = timed(args.steps, w.iterate)
    launches = reduce_sum(float(sip.kernel_launches() - launches0))
    clocks = sampler.stop() if sampler else None
    w.before_launch = w.after_launch = None
    k_ms = sum(a.elapsed_time(b) for _, a, b in kev)
    k_flops = sum(f for f, _, _ in kev)
    k_tf = reduce_sum(k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0) / world   # per-GPU achieved, mean over ranks
    k_share = reduce_max(k_ms) / ms_total if ms_total else None
    ms_step = ms_total / args.steps
    value = flops / (ms_step * 1e-3) / 1e12

    # ---- timed region 2: end to end through the public API with HOST buffers (pinned) ----
    e2e = None
    if not args.no_e2e:
        nbytes = w.T2old.local_bytes()
        h_in = api.lib().sipgpu_host_alloc(max(nbytes, 8))
        h_out = api.lib().sipgpu_host_alloc(max(nbytes, 8))
        api._check(api.lib().sipgpu_d2h(h_in, w.T2old.local_base(), nbytes // 8), "d2h")  # the host copy of T2old

        def e2e_step():
            w.load_t2old_from_host(h_in)
            e = w.iterate()
            w.store_t2new_to_host(h_out)
            return e

        # The kernels are warm from region 1 and the copies go through pinned buffers, so no extra untimed pass; at most
        # two end-to-end steps are timed (36 s each at full size) so that a large --steps stays bounded.
        e2e_steps = max(1, min(args.steps, 2))
        ms_e2e, energy_e2e = timed(e2e_steps, e2e_step)
        assert abs(energy_e2e - energy) <= 1e-9 * max(1.0, abs(energy)), (energy_e2e, energy)
        e2e = {"value": flops / (ms_e2e / e2e_steps * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": int(reduce_sum(float(nbytes))), "d2h_bytes_per_step": int(reduce_sum(float(nbytes + 8))),
               "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps}
        api.lib().sipgpu_host_free(h_in)
        api.lib().sipgpu_host_free(h_out)

    # ---- block traffic over NVLink ("NVLink GB/s for get/put"): every rank fetches / accumulates the blocks
    # owned by its neighbour, outside the timed regions; GB/s per GPU, mean over ranks ----
    nvlink = None
    if world > 1:
        peer_blocks = [b for b in w.blocks if w.T2old.owner(b) == (rank + 1) % world][:96]
        nbytes_blk = 8.0 * int(np.prod(w.T2old.block_shape(peer_blocks[0])))
        tmp_blocks = [api.DeviceBlock(w.T2old.block_shape(b)) for b in peer_blocks]

        def do_get():
            for b, t in zip(peer_blocks, tmp_blocks):
                w.T2old.get(b, out=t)

        def do_put_acc():
            for b, t in zip(peer_blocks, tmp_blocks):
                w.Xs.put_accumulate(b, t)

        do_get()
        ms_get, _ = timed(1, do_get)
        for t in tmp_blocks:
            t.fill(0.0)          # adds zeros: Xs is scratch that the next iteration overwrites anyway
        do_put_acc()
        ms_put, _ = timed(1, do_put_acc)
        vol = len(peer_blocks) * nbytes_blk
        nvlink = {"get_GBps_per_gpu": vol / (ms_get * 1e-3) / 1e9, "put_accumulate_GBps_per_gpu": vol / (ms_put * 1e-3) / 1e9,
                  "blocks": len(peer_blocks), "block_bytes": nbytes_blk,
                  "how": "peer reads (cudaMemcpyAsync from the owner's IPC-mapped slab) / red.global.add.f64 into the owner's "
                         "slab, all ranks concurrently towards rank+1, max over ranks of the elapsed time"}
        del tmp_blocks

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        s = CpuSample(o_segs, v_segs, args.cpu_dests, cores)
        s.run_once()
        sec, _ = s.run_once()
        cpu = {"value": s.flops / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
               "sample": f"{len(s.dests)} destination blocks x all 5 terms ({s.flops / 1e9:.0f} GFLOP, {sec:.1f} s), oracle "
                         f"permute->OpenBLAS dgemm->permute + accumulate"}

    # DRAM traffic of the dominant launch (the pp-ladder chain launch, ~76 % of the step) from the committed single-pass
    # ncu capture of this same command at full size; null for development sizes
    traffic = traffic_alg = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic_ccsd_full.json")
    if o_segs == O_SEGS and v_segs == V_SEGS and args.density >= 1.0 and os.path.exists(tpath):
        try:
            dom = max(json.load(open(tpath))["launches"], key=lambda x: x["seconds"])
            traffic, traffic_alg = dom["dram_bytes"] / world, dom["algorithmic_min_bytes"] / world
        except Exception:
            traffic = traffic_alg = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(o_segs, v_segs, args.density), "flops_per_step": flops,
                       "parallelism": f"owner-computes over {world} GPU(s), destination blocks block-cyclic",
                       "l2": "operand arrays (>= 10 GB each) exceed the 126 MB L2; no flush needed",
                       "energy": energy},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "contract_kernel (FP64 DMMA, mma.sync.m8n8k4.f64)",
                         "achieved": k_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": k_tf / peak_tf if peak_tf else None,
                         "peak_source": "DMMA issue-rate probe measured in this run (MEASURED_PEAKS.json holds no FP64 figure)",
                         "kernel_share_of_step": k_share,
                         "traffic": traffic, "traffic_algorithmic_min": traffic_alg,
                         "traffic_source": "profiles/r01_traffic_ccsd_full.json (ncu dram__bytes_read+write of the dominant "
                                           "launch, the pp-ladder chain; bytes per launch per GPU)" if traffic else None},
            "cpu_baseline": cpu,
            "nvlink": nvlink,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
