"""bench.py's reference arm (`--impl reference`: the CPU restatement of the reference's path on the host cores) at toy
sizes -- the one leg of bench.py that runs without a GPU -- checked against the JSON contract the driver reads: one line,
the base keys, `impl`, a `cpu_baseline` describing this very run and an `e2e` that moves no bytes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--o-segs", "4,4", "--v-segs",
                          "6,6,6", "--steps", "2", "--warmup", "1", "--cpu-dests", "1", *extra], capture_output=True, text=True,
                         timeout=300, env=e)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_contract_line():
    lines = [ln for ln in run().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    # steps / warmup are the ACTUAL counts (a wall budget caps them at full size; the first-touch pass and the thread-count
    # probes are warm-up passes); what was asked for is reported next to them
    assert d["steps"] == 2 and d["steps_requested"] == 2 and d["warmup"] >= 1 and d["warmup_requested"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert set(d["config"]) == {"workload", "flops_per_step", "parallelism", "l2"}    # same keys as the GPU arm's config
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "destination blocks" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_under_torchrun_only_rank_zero_speaks():
    assert run(env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""
    assert json.loads(run("--gpus", "2", env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}))["n_gpus"] == 2
