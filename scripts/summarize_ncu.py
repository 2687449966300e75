"""Turn ncu outputs into the compact text summaries committed under profiles/.

  python scripts/summarize_ncu.py rep  <file.ncu-rep> > profiles/<name>.txt     # one --set full capture
  python scripts/summarize_ncu.py list <launches.csv> > profiles/<name>.txt     # a --metrics gpu__time_duration.sum list
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def rep(path, top=14):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    # split the source page per kernel
    blocks, cur = [], None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) > 10:
            cur["rows"].append(r)
    print(f"# ncu --set full summary of {path.split('/')[-1]} (scripts/summarize_ncu.py; times under the profiler, clock-control none)")
    for k, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"\n## launch {k}: {d.get('Kernel Name', '?')}")
        for key in KEYS:
            if key in d:
                print(f"  {key:88s} {d[key]:>16s} {u[key]}")
        st = [(h, float(d[h])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h
              and d[h] not in ("", "n/a")]
        st.sort(key=lambda x: -x[1])
        print("  warp stall reasons (cycles per issued instruction, top 6): " +
              ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for h, v in st[:6]))
        if k < len(blocks) and blocks[k]["hdr"]:
            ix = {h: i for i, h in enumerate(blocks[k]["hdr"])}
            data = blocks[k]["rows"]
            stalls = [h for h in blocks[k]["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
            tot = sum(int(x[ix["# Samples"]]) for x in data)
            print(f"  pc sampling: {tot} samples over {len(data)} SASS instructions; share by stall reason: " +
                  ", ".join(f"{s_[6:]} {100.0 * sum(int(x[ix[s_]]) for x in data) / max(tot, 1):.1f}%" for s_ in stalls
                            if sum(int(x[ix[s_]]) for x in data) > 0.005 * tot))
            mn = OrderedDict()
            for x in data:
                op = x[ix["Source"]].split()[0] if x[ix["Source"]].split() else "?"
                if op.startswith("@"):
                    op = x[ix["Source"]].split()[1]
                op = op.split(".")[0]
                mn[op] = mn.get(op, 0) + int(x[ix["Instructions Executed"]] or 0)
            tops = sorted(mn.items(), key=lambda kv: -kv[1])[:top]
            print("  executed warp instructions by opcode: " + ", ".join(f"{o} {c}" for o, c in tops))


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        unit = r[ix["Metric Unit"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list {path.split('/')[-1]}: per-kernel totals (cold-cache, serialised: compare SHARES)")
    print(f"# {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms total")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * us / tot:6.2f}%  {us / 1e3:12.3f} ms  {n:7d} launches  {name}")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
