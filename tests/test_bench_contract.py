"""CPU: the bench line contract.  Checks the JSON lines committed from the last GPU runs of this round
(profiles/r2_bench_line.json and r2_bench_reference_line.json, written by `python bench.py [--impl reference]` on a
B200) for every key the driver reads, the scaling lines for their per-rank figures, and that bench.py parses its
arguments and maps device ordinals without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_every_contract_key():
    b = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line.json")))
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
    # round 2: per-rank times, the strong-scaling split, the streams sweep, where the clocks came from
    assert len(b["per_rank_ms"]) == b["n_gpus"] and abs(max(b["per_rank_ms"]) - b["ms_per_step"]) < 1e-9
    assert b["strong_scaling"]["streams_total"] == 256 and sum(b["strong_scaling"]["streams_per_gpu"]) == 256
    assert [p["streams"] for p in b["streams_sweep"]] == [32, 64, 128, 256, 512, 1024]
    assert b["clocks"]["source"] in ("nvml", "nvidia-smi -lms") and b["clocks"]["sm_mhz_min"] <= b["clocks"]["sm_mhz"]
    assert b["clocks"]["per_gpu_timed_steps"] and b["clocks"]["per_gpu_timed_steps"][0]["sm_mhz_min"] > 0.9 * b["clocks"]["sm_max_mhz"]
    assert b["roofline"]["traffic"] is not None and "sass" in b["roofline"]["traffic_source"].lower()


def test_committed_reference_arm_line():
    r = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_reference_line.json")))
    b = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line.json")))
    assert r["impl"] == "reference" and r["metric"] == b["metric"] and r["unit"] == b["unit"]
    assert r["higher_is_better"] is True and r["config"]["workload"] == b["config"]["workload"]
    assert r["cpu_baseline"]["kind"] in ("reference", "port") and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"]["value"] == r["value"] and r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["gpu_launches"] == 0


def test_committed_scaling_lines():
    """weak scaling from the per-rank figures of the committed 1 / 2 / 4 / 8 GPU lines: every rank within 2 % of the
    fastest once every rank processes the same kind of streams, and the N = 8 efficiency the summary quotes"""
    v = {}
    for n in (1, 2, 4, 8):
        d = json.load(open(os.path.join(ROOT, "profiles", f"r2_scale_n{n}.json")))
        assert d["n_gpus"] == n and len(d["per_rank_ms"]) == n and d["scaling"] == "weak"
        assert max(d["per_rank_ms"]) < 1.02 * min(d["per_rank_ms"])
        assert sum(d["strong_scaling"]["streams_per_gpu"]) == 256 and len(d["strong_scaling"]["per_rank_ms"]) == n
        v[n] = d["value"]
    assert v[8] / (8 * v[1]) > 0.97 and v[4] / (4 * v[1]) > 0.97 and v[2] / (2 * v[1]) > 0.97
    # the run that showed the cause: the slowest ranks stay ranks 6 and 7 whichever GPU they drive
    for f in ("r2_scale_n8_tone_in_pilot.json", "r2_scale_n8_tone_in_pilot_gpus_reversed.json"):
        t = json.load(open(os.path.join(ROOT, "profiles", f)))["per_rank_ms"]
        assert sorted(range(8), key=lambda i: t[i])[-2:] == [6, 7]


def test_clock_sampler_maps_cuda_ordinals_to_boards():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.visible_to_physical([0], {"CUDA_VISIBLE_DEVICES": "3"}) == ["3"]
    assert bench.visible_to_physical(range(3), {"CUDA_VISIBLE_DEVICES": "7,6,5,4"}) == ["7", "6", "5"]
    assert bench.visible_to_physical([0, 1], {}) == ["0", "1"]
    assert bench.visible_to_physical([0], {"CUDA_VISIBLE_DEVICES": "GPU-1234abcd"}) == ["GPU-1234abcd"]
    c = bench.ClockSampler([0])                   # no NVML / nvidia-smi here: the sampler must still start, stop and summarise
    with c:
        c.window_begin()
        c.window_end()
    s = c.summary()
    assert s["samples"] == 0 and s["reasons"] == [] and s["per_gpu"] == [] and s["per_gpu_timed_steps"] == []


def test_bench_parses_arguments_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in r.stdout + r.stderr                       # bench.py keeps stdout for the JSON line
