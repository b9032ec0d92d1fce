"""CPU: the bench.py contract that can be checked without a GPU -- the `--impl reference` arm (the oracle
port timed on the host cores) prints ONE JSON line with the keys the driver reads, on the same metric /
unit / config object as this repo's arm, and non-zero ranks of a multi-rank launch exit without work."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                         env=e, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    return [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_and_config():
    sys.path.insert(0, ROOT)
    import bench

    lines = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - bench.WORKLOAD["N"]) < 1e-3 * bench.WORKLOAD["N"]
    # the same config object as this repo's arm at the same N
    assert d["config"] == bench.workload_config(2)
    assert d["config"]["workload"].startswith("configs[1]")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    assert run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                     env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
