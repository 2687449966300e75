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
    ap.add_argument("--wall-budget", type=float, default=None,
                    help="seconds this process may take in total (default 600, reference arm 150): a full-size iteration is "
                         "36 s on one B200, so --steps/--warmup are capped to what fits (never below 3 warm-ups / 2 timed "
                         "steps unless fewer were asked for); the line reports the ACTUAL steps/warmup and *_requested")
    ap.add_argument("--verify-blocks", type=int, default=None,
                    help="destination blocks per rank recomputed locally and compared with the owner's copy (default: 64 at "
                         "N > 1, 8 at N = 1)")
    return ap.parse_args()


T_START = time.monotonic()   # the wall budget counts from process start (imports included)
# energy functional of the default workload (o = 3x20, v = 12x50, density 1, seed 0xACE54) from the N = 1 run of record
# (BENCH_r01.json): every N must reproduce it -- the iteration's result does not depend on the partition
ENERGY_N1 = 4286482.921305567


def config_of(o_segs, v_segs, density, world):
    """`config` is the same object in both arms (the driver compares them)."""
    return {"workload": workload_name(o_segs, v_segs, density),
            "flops_per_step": None,   # filled by the caller: algorithmic flops of one whole iteration
            "parallelism": f"destination blocks block-cyclic over {world} rank(s), owner computes",
            "l2": "operand arrays (>= 10 GB each at full size) exceed the 126 MB L2; no flush needed"}


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
        self.threads = threads
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

    def set_threads(self, n):
        """Both thread pools of the CPU arm: OpenBLAS (dgemm) and libgomp (the oracle's permute loops)."""
        import ctypes
        self.oracle.use_openblas._keep.scipy_openblas_set_num_threads(int(n))
        try:
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
        except OSError:
            pass
        self.threads = int(n)

    def pick_threads(self, cores):
        """More threads are not always faster here (SCALE_r01: 32 cores slower than 16 -- NUMA / oversubscription of the
        memory-bound permutes): time one pass at 8 (the reference's own OpenMP cap, src/sip/core/sip.cpp:113), 16 and all
        cores, keep the fastest.  Returns {threads: seconds}."""
        seen = {}
        for n in sorted({min(8, cores), min(16, cores), cores}):
            self.set_threads(n)
            seen[n] = self.run_once()[0]
        self.set_threads(min(seen, key=seen.get))
        return seen

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
    from aces4_b200.sial_workload import iteration_flops
    budget = args.wall_budget if args.wall_budget is not None else 150.0
    cores = host_cores()
    s = CpuSample(o_segs, v_segs, args.cpu_dests, cores)
    s.run_once()                       # first touch of every operand block (page faults), not a measurement
    seen = s.pick_threads(cores)       # counts as the warm-up passes
    per = seen[s.threads]
    left = budget - (time.monotonic() - T_START) - 5.0
    warm = max(0, min(args.warmup - len(seen), int(left / per) - min(args.steps, 3)))
    for _ in range(warm):
        s.run_once()
    left = budget - (time.monotonic() - T_START) - 5.0
    steps = max(1, min(args.steps, max(min(args.steps, 3), int(left / per))))
    times = [s.run_once()[0] for _ in range(steps)]
    sec = sum(times) / len(times)
    tf = s.flops / sec / 1e12
    sample = (f"{len(s.dests)} of {len(v_segs) ** 2 * len(o_segs) ** 2} destination blocks x all 5 terms "
              f"({s.flops / 1e9:.0f} GFLOP per step), oracle permute->dgemm->permute + accumulate, {s.blas.split()[0]} dgemm, "
              f"{s.threads} threads (fastest of " + ", ".join(f"{n}: {t:.1f} s" for n, t in sorted(seen.items())) + ")")
    cfg = config_of(o_segs, v_segs, args.density, args.gpus)
    cfg["flops_per_step"] = iteration_flops(o_segs, v_segs)
    line = {"metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm + len(seen) + 1, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "wall_budget_s": budget, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": cfg,
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": s.threads, "host_cores": cores, "kind": "port",
                             "sample": sample, "threads_tried_s": {str(k): v for k, v in seen.items()}},
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

    def timed(nsteps, body, label=None):
        """(max-over-ranks ms of the whole region, last result).  One event pair brackets the region; when `label` is
        given a provisional line goes to stderr after every step, so a run that is killed leaves its progress behind."""
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = None
        for k in range(nsteps):
            res = body()
            if label and rank == 0:
                ek = torch.cuda.Event(enable_timing=True)
                ek.record(stream)
                ek.synchronize()   # body() ended with a blocking scalar read-back: nothing is pending
                ms = e0.elapsed_time(ek) / (k + 1)
                print(json.dumps({"provisional": label, "steps_done": k + 1, "ms_per_step": ms,
                                  "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s"}), file=sys.stderr, flush=True)
        e1.record(stream)
        e1.synchronize()
        fence()
        return reduce_max(e0.elapsed_time(e1)), res

    # ---- warm-up and the step plan: everything has to fit the wall budget (a full-size step is 36 s at N = 1) ----
    budget = args.wall_budget if args.wall_budget is not None else 600.0
    t0 = time.monotonic()
    energy = w.iterate()          # warm-up 1: also the measure of a step's wall time
    sip.sync()
    step_s = reduce_max(time.monotonic() - t0)
    elapsed = reduce_max(time.monotonic() - T_START)
    nverify = args.verify_blocks if args.verify_blocks is not None else (64 if world > 1 else 8)
    reserve = 15.0 + (0.0 if args.no_e2e else 1.1 * step_s) + nverify * 0.05 * min(1.0, step_s)
    if world == 1 and not args.no_cpu_baseline:
        reserve += 45.0           # CPU leg: one first-touch pass + two passes at 8 threads and at all cores
    n_avail = int(max(0.0, budget - elapsed - reserve) / step_s)
    if (args.warmup - 1) + args.steps <= n_avail:
        warm, steps = args.warmup, args.steps
    else:
        warm = min(args.warmup, 3)                                        # timing rule: W >= 3 whenever it was asked for
        steps = min(args.steps, max(min(args.steps, 2), n_avail - (warm - 1)))
    for _ in range(max(0, warm - 1)):
        energy = w.iterate()
    warm = max(warm, 1)

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
    ms_total, energy = timed(steps, w.iterate, label="resident")
    launches = reduce_sum(float(sip.kernel_launches() - launches0))
    clocks = sampler.stop() if sampler else None
    w.before_launch = w.after_launch = None
    k_ms = sum(a.elapsed_time(b) for _, a, b in kev)
    k_flops = sum(f for f, _, _ in kev)
    k_tf = reduce_sum(k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0) / world   # per-GPU achieved, mean over ranks
    k_share = reduce_max(k_ms) / ms_total if ms_total else None
    ms_step = ms_total / steps
    value = flops / (ms_step * 1e-3) / 1e12

    # ---- parity of the distributed result: the energy functional against the N = 1 value of record, and sampled T2new
    # blocks (fetched from their owners: peer reads) against the same blocks recomputed on this rank from the replicated
    # inputs alone.  A wrong put += into peer HBM or a wrong peer get shows up here, in the bench line itself. ----
    parity = None
    if nverify > 0:
        sip.sync()
        if world > 1:
            barrier()
        worst = reduce_max(w.verify_blocks(nverify, offset=rank * 17))
        parity = {"blocks_checked_per_rank": nverify, "t2new_max_rel_err": worst, "tolerance": 1e-10}
        if o_segs == O_SEGS and v_segs == V_SEGS and args.density >= 1.0:
            parity["energy"] = energy
            parity["energy_n1"] = ENERGY_N1
            parity["energy_rel_err"] = abs(energy - ENERGY_N1) / abs(ENERGY_N1)
            parity["ok"] = bool(worst <= 1e-10 and parity["energy_rel_err"] <= 1e-12)
        else:
            parity["ok"] = bool(worst <= 1e-10)
        if world > 1:
            barrier()
        if not parity["ok"]:
            raise SystemExit(f"parity check failed: {json.dumps(parity)}")

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

        # The kernels are warm from region 1 and the copies go through pinned buffers, so no extra untimed pass; the
        # number of end-to-end steps is what is left of the wall budget (at least 1, at most what region 1 timed).
        left = budget - reduce_max(time.monotonic() - T_START) - (reserve - 1.1 * step_s)
        e2e_steps = max(1, min(steps, 3, int(left / (1.05 * step_s))))   # at most 3: the figure is stable after one
        ms_e2e, energy_e2e = timed(e2e_steps, e2e_step, label="e2e")
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

        sec_get = w.T2old.section(peer_blocks, tmp_blocks)
        sec_put = w.Xs.section(peer_blocks, tmp_blocks)

        def do_get():        # one section's worth of gets in one C-ABI call: ONE gather launch over the peer-mapped slabs
            w.T2old.get_many(section=sec_get)

        def do_put_acc():    # ONE red.global.add.f64 launch into the neighbour's slab
            w.Xs.put_accumulate_many(section=sec_put)

        def do_get_eager():
            for b, t in zip(peer_blocks, tmp_blocks):
                w.T2old.get(b, out=t)

        do_get()
        ms_get, _ = timed(1, do_get)
        do_get_eager()
        ms_get_eager, _ = timed(1, do_get_eager)
        for t in tmp_blocks:
            t.fill(0.0)          # adds zeros: Xs is scratch that the next iteration overwrites anyway
        do_put_acc()
        ms_put, _ = timed(1, do_put_acc)
        vol = len(peer_blocks) * nbytes_blk
        nvlink = {"get_GBps_per_gpu": vol / (ms_get * 1e-3) / 1e9, "put_accumulate_GBps_per_gpu": vol / (ms_put * 1e-3) / 1e9,
                  "get_per_block_memcpy_GBps_per_gpu": vol / (ms_get_eager * 1e-3) / 1e9,
                  "blocks": len(peer_blocks), "block_bytes": nbytes_blk,
                  "how": "one section of gets / put += towards rank+1 through sipgpu_array_get_many / _put_accumulate_many: one descriptor-driven launch "
                         "of TMA bulk transfers through a shared-memory ring: cp.async.bulk from the owner's IPC-mapped slab (get) / "
                         "cp.reduce.async.bulk add.f64 into it (put +=); all ranks concurrently (two-way traffic: the wire's ceiling "
                         "with 128 B packets is 686 / 720 GB/s, profiles/r02_nvlink_counters.txt), max over ranks of the elapsed time; per_block_memcpy = the same gets as one "
                         "cudaMemcpyAsync per block"}
        del tmp_blocks

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        s = CpuSample(o_segs, v_segs, args.cpu_dests, cores)
        s.run_once()                     # first touch
        seen = s.pick_threads(cores)
        sec = seen[s.threads]
        cpu = {"value": s.flops / sec / 1e12, "unit": "TFLOP/s", "cores": s.threads, "host_cores": cores, "kind": "port",
               "sample": f"{len(s.dests)} destination blocks x all 5 terms ({s.flops / 1e9:.0f} GFLOP, {sec:.1f} s), oracle "
                         f"permute->OpenBLAS dgemm->permute + accumulate; fastest of the thread counts tried",
               "threads_tried_TFLOPs": {str(n): s.flops / t / 1e12 for n, t in sorted(seen.items())},
               "omp8": {"value": s.flops / seen[min(8, cores)] / 1e12, "unit": "TFLOP/s",
                        "note": "8 threads = the reference's own OpenMP cap (src/sip/core/sip.cpp:113)"}}

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
        cfg = config_of(o_segs, v_segs, args.density, world)
        cfg["flops_per_step"] = flops
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup, "wall_budget_s": budget,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "energy": energy, "parity_vs_n1": parity,
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
            "wall_s": time.monotonic() - T_START,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
