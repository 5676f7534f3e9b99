"""CPU: the bench line contract.  Checks the JSON line committed from the last GPU run of this round
(profiles/r1_bench_line.json, written by `python bench.py` on a B200) for every key the driver reads, and
that bench.py parses its arguments without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_every_contract_key():
    b = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_line.json")))
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e",
              "gpu_launches", "clocks"):
        assert k in b, k
    assert b["n_gpus"] == 1 and b["warmup"] >= 3 and b["higher_is_better"] is True and b["scaling"] == "weak"
    assert b["vs_baseline"] is None                                  # BASELINE.md publishes no number for this metric
    assert "workload" in b["config"] and "model" not in b["config"]
    assert "MS/s" in b["unit"] and isinstance(base.get("metric", ""), str)
    r = b["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"] * 0.99
    c = b["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = b["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < b["value"]
    assert b["gpu_launches"] > 0
    assert not set(b["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # consistency of the line with itself: value = samples per step / time per step
    samples = b["config"]["streams_per_gpu"] * b["config"]["samples_per_stream"]
    assert abs(b["value"] - samples / b["ms_per_step"] / 1e3) < 1e-6 * b["value"]


def test_bench_parses_arguments_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in r.stdout + r.stderr                       # bench.py keeps stdout for the JSON line
