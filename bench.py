#!/usr/bin/env python
"""bench.py — IQ MS/s through the FM demodulation chain on N B200s (one rank per GPU).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON
line (rank 0).  A "step" = one pass of the hot path over one batch of synthetic IQ:
`--streams` independent 2.304 MS/s streams x `--seconds` of signal PER GPU (weak scaling:
the batch per GPU is fixed, streams are sharded over ranks, there is no data-path
collective; the only collective is the one-time broadcast of the tap/LUT blob).

  value     whole-job throughput, inputs resident in HBM, CUDA-event timed on the handle's
            stream, max over ranks
  e2e       same metric through the host-buffer C-ABI call (sdrjfm_process): pinned host
            IQ in, audio + RDS baseband out, copies inside the timed region
  roofline  the decimating front-end kernel timed alone: 8.667 algorithmic bytes per input
            sample (8 B float2 read + 8/12 B fm-rate float2 write)
  cpu_baseline  the reference's own DSP classes (oracle/_ref) — or the port when _ref is
            absent — on the host cores, bounded sample (rank 0, N=1 only): all cores, one stream
            per thread, plus the single-core figure
  strong_scaling  BASELINE config 5 as written: `--streams` streams IN TOTAL sharded over the N
            ranks (sharding.stream_range), same bytes per step as the N=1 line
  streams_sweep   (N=1) the chain at 32 .. 1024 streams per GPU at constant bytes per step
  per_rank_ms     ms per step of every rank (value uses the max)
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# 11 CUDA streams per handle (lanes + side streams): more hardware queues than the default 8; must be in
# the environment before the CUDA context exists (torch creates it first in this process)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from __graft_entry__ import load_package  # noqa: E402

INPUT_RATE = 2304000
ALGO_BYTES_PER_SAMPLE = 8.0 + 8.0 / 12.0
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md, used if MEASURED_PEAKS.json is absent


def chain_settings():
    """settings of the benchmarked chain = BASELINE.json config 5 (stereo + pilot/PSS + RDS)
    as far as the GPU path implements it; the same dict configures the CPU reference arm."""
    return dict(fm_mode=0, decoder=3, rds_on=1, auto_mono=1, pss_on=1, dc_remove=1,
                deemph_us=50, volume_db=-6.0)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
                if k in d:
                    return float(d[k]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def kernel_sass_hash(kernel="frontend_tma_kernelILi12ELi4ELi37"):
    """sha256 of the SASS of one kernel of the loaded library (addresses and encodings stripped): the committed
    ncu traffic figure is only reported while the kernel it was captured from is the kernel that runs."""
    import hashlib
    import re
    so = os.path.join(ROOT, "sdr-j-fm_b200", "libsdrjfm_b200.so")
    try:
        out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    keep, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = kernel in ln
            continue
        if on:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
            if m:
                keep.append(m.group(1).strip())
    if not keep:
        return None
    return hashlib.sha256("\n".join(keep).encode()).hexdigest()[:16]


def audio48_parity():
    """the reference's 192 -> 48 kHz step is libsamplerate (newconverter.cpp:37,55-80): pinned only where the
    library exists (tests/test_gpu_parity.py::test_audio48_against_libsamplerate_if_present)"""
    import ctypes.util
    name = ctypes.util.find_library("samplerate")
    return f"libsamplerate present ({name}): see the GPU test log" if name else "unpinned (libsamplerate absent)"


def host_topology(index):
    """NUMA placement of this rank: the CPUs NVML names for the GPU, the node they sit on, nodes of the box"""
    info = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = sorted(64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1)
        info["gpu_cpu_affinity"] = f"{cpus[0]}-{cpus[-1]}" if cpus else None
    except Exception:
        pass
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        info["numa_nodes"] = len(nodes)
        info["numa_cpulists"] = {d: open(f"/sys/devices/system/node/{d}/cpulist").read().strip() for d in nodes}
    except Exception:
        pass
    info["process_cpus"] = len(os.sched_getaffinity(0))
    return info


def visible_to_physical(indices, env=None):
    """CUDA device ordinals of this process -> what NVML / nvidia-smi call the same boards: the entries of
    CUDA_VISIBLE_DEVICES (physical indices or GPU-... UUIDs) when it is set, the ordinals themselves otherwise."""
    cvd = (os.environ if env is None else env).get("CUDA_VISIBLE_DEVICES", "")
    toks = [t.strip() for t in cvd.split(",") if t.strip()]
    return [toks[i] if i < len(toks) else str(i) for i in indices] if toks else [str(i) for i in indices]


class ClockSampler:
    """samples SM clocks / throttle reasons during the timed region: through NVML in this process (what nvidia-smi
    reads; no process is spawned while the GPUs are being timed), else with the recipe's looping nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason bits (nvml.h): SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
    BITS = (0x8, 0x40, 0x20, 0x4)

    def __init__(self, indices, period=None):
        """indices: the GPUs of this job.  ONE sampler per job (rank 0) reads all of them: eight ranks each starting
        an nvidia-smi five times a second showed as a 10 % slower slowest rank at N=8, one rank doing so as 7 %
        (every new nvidia-smi attaches to all eight GPUs while they are timed; profiles/r2_summary.md)."""
        self.indices, self.rows, self.stop, self.period = list(indices), [], threading.Event(), period
        self.win = None
        self.source, self.nvml, self.handles, self.proc = "nvidia-smi -lms", None, [], None
        self.per_gpu = {}
        # CUDA ordinals -> the board NVML / nvidia-smi must be asked about (CUDA_VISIBLE_DEVICES may renumber or name them)
        self.ids = visible_to_physical(self.indices)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(int(t)) if t.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(t)
                            for t in self.ids]
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None
        if self.period is None:
            self.period = 0.05 if self.nvml is not None else 0.2
        self.t = threading.Thread(target=self.run, daemon=True)

    def window_begin(self):
        """the device-resident timed steps (`value`): per-GPU figures are kept for this window alone as well"""
        self.win = {gi: len(self.per_gpu.get(gi, {"sm": []})["sm"]) for gi in self.indices}

    def window_end(self):
        if self.win is not None:
            # (a pass over eight GPUs takes NVML ~0.1 s under load: the samples next to the window count as well)
            self.win = {gi: (max(a - 1, 0), len(self.per_gpu.get(gi, {"sm": []})["sm"]) + 1) for gi, a in self.win.items()}

    def run(self):
        if self.nvml is None:
            # the recipe's form: ONE nvidia-smi that stays attached and prints a row per GPU every period
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(self.ids),
                                              f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-lms", str(int(self.period * 1000))], stdout=subprocess.PIPE, text=True)
                for ln in self.proc.stdout:
                    if ln.strip():
                        self.rows.append([c.strip() for c in ln.split(",")])
                    if self.stop.is_set():
                        break
            except Exception:
                pass
            return
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop.is_set():
            for gi, h in zip(self.indices, self.handles):
                try:
                    r = int(reasons(h))
                    sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                    self.rows.append([str(sm), str(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))] +
                                     ["Active" if r & b else "Not Active" for b in self.BITS])
                    g = self.per_gpu.setdefault(gi, {"sm": [], "mem": [], "w": [], "reasons": 0})
                    g["sm"].append(sm)
                    g["mem"].append(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_MEM))
                    g["w"].append(n.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    g["reasons"] |= r
                except Exception:
                    pass
            self.stop.wait(self.period)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.t.join(timeout=6)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows), "sm_mhz_min": min(sm) if sm else None, "source": self.source,
                # per GPU of the job: median / min SM clock, min memory clock, median / max board power, reason bits seen
                "per_gpu": [{"gpu": gi, "board": dict(zip(self.indices, self.ids)).get(gi), "sm_mhz": int(np.median(g["sm"])), "sm_mhz_min": int(min(g["sm"])),
                             "mem_mhz_min": int(min(g["mem"])), "power_w": round(float(np.median(g["w"])), 1),
                             "power_w_max": round(float(max(g["w"])), 1), "reason_bits": hex(g["reasons"])}
                            for gi, g in sorted(self.per_gpu.items()) if g["sm"]],
                # the same for the samples taken inside the device-resident timed steps only
                "per_gpu_timed_steps": [{"gpu": gi, "samples": len(self.per_gpu[gi]["sm"][a:b]),
                                         "sm_mhz_min": int(min(self.per_gpu[gi]["sm"][a:b])),
                                         "mem_mhz_min": int(min(self.per_gpu[gi]["mem"][a:b])),
                                         "power_w": round(float(np.median(self.per_gpu[gi]["w"][a:b])), 1),
                                         "power_w_max": round(float(max(self.per_gpu[gi]["w"][a:b])), 1)}
                                        for gi, (a, b) in sorted((self.win or {}).items())
                                        if gi in self.per_gpu and self.per_gpu[gi]["sm"][a:b]]}


def gen_batch_gpu(torch, dev, n_streams, n, stream0=0, group=16):
    """config-5 style IQ on the GPU: stream s = stereo MPX, L tone 400 + 10 (s mod 256) Hz, 10 % pilot,
    57 kHz BPSK-like RDS sub-carrier, 75 kHz deviation, amp 0.5, AWGN 40 dB (seeded).
    The tone wraps at BASELINE config 5's 256 streams: with the global stream index in it, rank 7 of an 8-GPU
    weak-scaling run got audio tones of 18.3-20.9 kHz — on top of the 19 kHz pilot, where the pilot PLL solver needs
    more passes — and set the job's `value` 7 % below the other ranks (profiles/r2_summary.md, "N = 8").  Phase, noise
    and RDS bits stay per global stream."""
    out = torch.empty((n_streams, n), dtype=torch.complex64, device=dev)
    t = torch.arange(n, dtype=torch.float64, device=dev) / INPUT_RATE
    th = 2 * np.pi * 19000.0 * t
    sin_th, sin_2th, sin_3th = torch.sin(th), torch.sin(2 * th), torch.sin(3 * th)
    g = torch.Generator(device=dev)
    for s0 in range(0, n_streams, group):
        ss = torch.arange(s0, min(s0 + group, n_streams), device=dev, dtype=torch.float64)
        sid = ss + stream0
        L = torch.sin(2 * np.pi * (400.0 + 10.0 * torch.remainder(sid, 256.0))[:, None] * t[None, :])
        g.manual_seed(2000 + int(sid[0].item()))
        bits = torch.randint(0, 2, (len(ss), 4096), device=dev, generator=g)
        d = torch.cumsum(bits, dim=1) & 1
        bit_idx = torch.floor(t * 1187.5).long() % 4096
        half = torch.floor(t * 2375.0).long() & 1
        sym = (1.0 - 2.0 * d[:, bit_idx].double()) * (1.0 - 2.0 * half.double())[None, :]
        mpx = 0.45 * L + 0.45 * L * sin_2th[None, :] + 0.10 * sin_th[None, :] + 0.05 * sym * sin_3th[None, :]
        phi = torch.cumsum(2 * np.pi * 75000.0 / INPUT_RATE * mpx, dim=1) + 0.1 * sid[:, None]
        sigma = 0.5 / np.sqrt(2.0) * 10 ** (-40.0 / 20)
        noise = sigma * torch.randn((len(ss), n, 2), device=dev, generator=g, dtype=torch.float32)
        x = 0.5 * torch.polar(torch.ones_like(phi), phi)
        out[s0:s0 + len(ss)] = x.to(torch.complex64) + torch.view_as_complex(noise)
        del L, mpx, phi, x, noise, sym
    return out


def cpu_reference_rate(settings, seconds_per_thread, threads=None):
    """times the reference DSP classes (oracle/_ref; the port if absent) on host cores:
    one stream per thread, each `seconds_per_thread` (>= 10 s) of config-2/5 signal. Returns dict."""
    from oracle import chainlib
    chainlib.build(("oracle", "ref"))
    kind, prefix = ("reference", "ref") if chainlib.available("ref") else ("port", "orc")
    sig = importlib.import_module("sdrjfm_b200.signals")
    cores = threads or os.cpu_count() or 1
    block = sig.batch_stream(0, INPUT_RATE)           # 1 s block, fed repeatedly
    reps = max(1, int(round(seconds_per_thread)))
    chains = [chainlib.Chain(prefix, **settings) for _ in range(cores)]
    taps = ("audio192", "rds24")

    def work(c):
        for _ in range(reps):
            c.process(block, taps=taps)

    ths = [threading.Thread(target=work, args=(c,)) for c in chains]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = cores * reps * len(block)
    return {"value": total / dt / 1e6, "unit": "MS/s", "cores": cores, "kind": kind,
            "sample": f"{cores} streams x {reps} s of 2.304 MS/s stereo+pilot+RDS IQ, one stream per "
                      f"thread, the reference's classes up to the 192 kHz de-emphasised audio and the 24 kHz RDS "
                      f"baseband; EXCLUDES the 192->48 kHz libsamplerate step (library absent), which the GPU arm includes",
            "seconds": dt}


def cpu_reference_both(settings, seconds_per_thread):
    """all host cores (the headline CPU figure) and ONE core"""
    allc = cpu_reference_rate(settings, seconds_per_thread)
    one = cpu_reference_rate(settings, seconds_per_thread, threads=1)
    allc["single_core"] = {"value": one["value"], "unit": "MS/s", "cores": 1,
                           "x_realtime_2p304MSps": one["value"] / 2.304}
    return allc


def bind_to_gpu_numa_node(index):
    """one process per GPU: run (and first-touch the pinned host buffers) on the CPUs NVML names as
    closest to this GPU, so that the host->device copies of the ranks do not share one memory controller"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else a library prints to fd 1 during the
    run (e.g. NCCL's version banner) has been routed to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def front_end_sweep(pkg, torch, dev, local, S, peak):
    """the decimating front end alone (one launch over S streams, no lanes) for every rate / format:
    algorithmic bytes = bytes of one IQ sample in that format + 8 bytes per fm-rate output."""
    os.environ["SDRJFM_LANES"] = "1"
    out = []
    seconds = 0.25
    fmts = {"cf32": (torch.float32, 8), "s16": (torch.int16, 4), "u8": (torch.uint8, 2)}
    cases = [(2304000, 0, "cf32"), (2304000, 0, "s16"), (2304000, 0, "u8"), (2400000, 0, "u8"),
             (6000000, 0, "cf32"), (6000000, 0, "s16"), (10000000, 0, "cf32"), (10000000, 0, "s16"),
             (2400000, 1, "cf32"), (6000000, 1, "s16"), (10000000, 1, "s16")]
    try:
        for rate, mode, fmt in cases:
            dec = 5 if mode == 1 else pkg.front_end_decimation(rate)
            n = int(seconds * rate) // (dec * 512) * (dec * 512)
            p = pkg.FmProcessorB200(n_streams=S, input_rate=rate, max_samples_per_call=n, device=local,
                                    keep_taps=False, front_end_mode=mode)
            dt, bps = fmts[fmt]
            if fmt == "cf32":
                buf = torch.randn((S, n, 2), device=dev, dtype=torch.float32) * 0.3
            else:
                buf = torch.randint(0, 200, (S, n, 2), device=dev, dtype=dt)
            st = torch.cuda.ExternalStream(p.cuda_stream, device=dev)
            for _ in range(3):
                p.run_frontend_only_raw(buf.data_ptr(), fmt, 2048, n, n)
            p.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            with torch.cuda.stream(st):
                e0.record(st)
                for _ in range(reps):
                    p.run_frontend_only_raw(buf.data_ptr(), fmt, 2048, n, n)
                e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            algo = S * n * (bps + 8.0 / dec)
            out.append({"input_rate": rate, "front_end_mode": mode, "format": fmt, "decimation": dec,
                        "kernel_ms": ms, "GS_per_s": S * n / ms / 1e6, "achieved_GBps": algo / ms / 1e6,
                        "frac_of_hbm_peak": algo / ms / 1e6 / peak})
            p.close()
            del buf
    finally:
        del os.environ["SDRJFM_LANES"]
    return out


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=256, help="IQ streams per GPU")
    ap.add_argument("--seconds", type=float, default=0.5, help="signal seconds per stream per step")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="cpu baseline: signal seconds per thread")
    ap.add_argument("--no-sweep", action="store_true", help="skip the streams-per-GPU sweep (N=1) / strong-scaling leg (N>1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--rds-symbols", action="store_true",
                    help="also run the optional GPU RDS symbol stage (Costas + rdsDecoder_1 -> bits) in every step")
    ap.add_argument("--front-end-sweep", action="store_true",
                    help="also time the front-end kernel alone for every device rate / sample format "
                         "(BASELINE config 4 and SURVEY.md §8(f) rank 1); adds `front_end_sweep` to the line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = load_package()
    settings = chain_settings()
    n = int(args.seconds * INPUT_RATE) // 12 * 12
    workload = (f"batched multi-stream (BASELINE config 5 shape): {args.streams} streams/GPU x "
                f"{n} IQ samples @2.304 MS/s, stereo MPX + 19 kHz pilot + 57 kHz RDS sub-carrier")
    config = {"workload": workload, "streams_per_gpu": args.streams, "samples_per_stream": n,
              "settings": settings, "l2_policy": "inputs larger than L2 (per-step IQ batch >> 126 MB)",
              "chain": "DC-remove, FIR /12, Mixed discriminator, AFC, pilot PLL + lock, PSS, 38 kHz L-R "
                       "matrix, RDS band-pass/Hilbert/x3 mix//8, de-emphasis, gain, 192->48 kHz"}

    if args.impl == "reference":
        if rank != 0:
            return
        # one untimed warm-up pass (page-in, table design), then K timed steps of the bounded sample
        # (each step = cores x cpu_seconds of signal, all host threads), then the single-core figure
        cpu_reference_rate(settings, 1.0)
        r = cpu_reference_rate(settings, args.cpu_seconds)
        vals = [r["value"]]
        for _ in range(max(0, min(args.steps, 3) - 1)):
            vals.append(cpu_reference_rate(settings, args.cpu_seconds)["value"])
        v = float(np.mean(vals))
        one = cpu_reference_rate(settings, args.cpu_seconds, threads=1)
        line = {"impl": "reference", "metric": "IQ MS/s through full FM demod chain", "value": v,
                "unit": "MS/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": 1,
                "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "MS/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"],
                                 "single_core": {"value": one["value"], "unit": "MS/s", "cores": 1,
                                                 "x_realtime_2p304MSps": one["value"] / 2.304}},
                "audio48_parity": audio48_parity(),
                "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    bind_to_gpu_numa_node(local)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FM path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S = args.streams
    proc = pkg.FmProcessorB200(n_streams=S, max_samples_per_call=n, device=local, keep_taps=False)
    proc.configure(**settings)
    if args.rds_symbols:
        proc.setRdsSymbolStage(True)
        config["rds_symbol_stage"] = "on (Costas loop + rdsDecoder_1, one lane per stream)"
    # shared tap/LUT blob: rank 0 designs, everyone imports what rank 0 broadcast (NCCL)
    if world > 1:
        sh = importlib.import_module("sdrjfm_b200.sharding")
        proc.tables_import(sh.broadcast_tables(proc.tables_export() if rank == 0 else None, dist, device=dev))

    x = gen_batch_gpu(torch, dev, S, n, stream0=rank * S)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(proc.cuda_stream, device=dev)
    d_audio = torch.empty((S, n // 48 + 16), dtype=torch.complex64, device=dev)
    d_rds = torch.empty((S, n // 96 + 16), dtype=torch.complex64, device=dev)

    def step_device():
        proc.process_device(x.data_ptr(), n, x.stride(0), d_audio.data_ptr(), d_audio.stride(0),
                            d_rds.data_ptr(), d_rds.stride(0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, st=None):
        st = st or ext
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(steps):
                fn()
            e1.record(st)
        barrier()
        ms = e0.elapsed_time(e1)
        timed.per_rank = [ms]
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            timed.per_rank = [float(v.item()) for v in allt]
            ms = max(timed.per_rank)
        return ms

    for _ in range(args.warmup):
        step_device()
    proc.sync()
    l0 = proc.launch_count
    # clocks / throttle reasons are sampled over ALL timed regions of this run (device-resident steps,
    # front-end kernel alone, end-to-end steps); the sampler is stopped before the CPU baseline leg
    clk = ClockSampler(range(world) if world > 1 else [local])
    if rank == 0:
        clk.__enter__()
        time.sleep(0.05)                  # (a first sample of every GPU before the steps start)
        clk.window_begin()
    ms = timed(step_device, args.steps)
    if rank == 0:
        clk.window_end()
    per_rank_ms = [v / args.steps for v in timed.per_rank]
    launches = proc.launch_count - l0
    value = world * S * n * args.steps / (ms * 1e-3) / 1e6

    # roofline: the front-end kernel alone, as ONE launch over all streams of this GPU (a handle whose
    # streams are not split into lanes; the lanes of `proc` would issue one launch per lane concurrently)
    saved = os.environ.get("SDRJFM_LANES")
    os.environ["SDRJFM_LANES"] = "1"
    proc1 = pkg.FmProcessorB200(n_streams=S, max_samples_per_call=n, device=local, keep_taps=False)
    if saved is None:
        del os.environ["SDRJFM_LANES"]
    else:
        os.environ["SDRJFM_LANES"] = saved
    proc1.configure(**settings)
    ext1 = torch.cuda.ExternalStream(proc1.cuda_stream, device=dev)
    for _ in range(3):
        proc1.run_frontend_only(x.data_ptr(), n, x.stride(0))
    proc1.sync()
    fe_steps = max(args.steps, 5)
    fe_l0 = proc1.launch_count
    fe_ms = timed(lambda: proc1.run_frontend_only(x.data_ptr(), n, x.stride(0)), fe_steps, ext1) / fe_steps
    fe_launches = (proc1.launch_count - fe_l0) / fe_steps
    proc1.close()
    peak, peak_kind = peak_hbm()
    achieved = ALGO_BYTES_PER_SAMPLE * S * n / (fe_ms * 1e-3) / 1e9
    # dram bytes per launch: from the committed ncu --set full capture of this workload — only while the kernel
    # that capture profiled is the kernel in the loaded library (SASS hash), else null
    traffic, traffic_note = None, "no capture of this kernel build / workload committed"
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_frontend_traffic.json")))
        if tr["streams"] == S and tr["samples_per_stream"] == n and tr["kernel"] == "frontend_tma_kernel":
            have = kernel_sass_hash() if rank == 0 else None
            if have is not None and have == tr.get("sass_sha16"):
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                traffic_note = f"ncu capture {tr.get('capture', '')}, kernel SASS sha {have}"
            else:
                traffic_note = f"committed capture is of another kernel build (sha {tr.get('sass_sha16')} vs loaded {have})"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "frontend_tma_kernel", "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_note,
                "kernel_ms": fe_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * S * n,
                "launches_timed": fe_launches,
                "note": "one cp.async.bulk.tensor-fed launch over all streams (+ one small launch for the "
                        "ragged last tile of each stream when n/12 is not a multiple of 512)"}

    # end to end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((S, n), dtype=torch.complex64, pin_memory=True)
        hx.copy_(x)
        ha = torch.empty((S, n // 48 + 16), dtype=torch.complex64, pin_memory=True)
        hr = torch.empty((S, n // 96 + 16), dtype=torch.complex64, pin_memory=True)
        import ctypes as C
        na, nr = C.c_int64(0), C.c_int64(0)

        def step_host():
            rc = proc.L.sdrjfm_process(proc.h, hx.data_ptr(), n, hx.stride(0), ha.data_ptr(),
                                       ha.stride(0), C.byref(na), hr.data_ptr(), hr.stride(0),
                                       C.byref(nr), None)
            assert rc == 0, proc.L.sdrjfm_last_error(proc.h)

        step_host()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            step_host()          # synchronous: returns after the D2H copies completed
        barrier()
        dt = time.perf_counter() - t0
        dts = [dt]
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            dts = [float(v.item()) for v in allt]
            dt = max(dts)
        e2e = {"value": world * S * n * e2e_steps / dt / 1e6, "unit": "MS/s",
               "h2d_bytes_per_step": S * n * 8,
               "d2h_bytes_per_step": S * (na.value + nr.value) * 8, "steps": e2e_steps,
               "h2d_GBps_per_rank": [S * n * 8 * e2e_steps / d / 1e9 for d in dts],
               "host": host_topology(local),
               "note": "cf32 (8 B/sample) is bound by the host->device copy: one PCIe Gen5 x16 link per GPU, and all "
                       "links of the box draw on the host memory of the NUMA node(s) listed; device_format_u8 is the "
                       "same stream at 2 B/sample"}

    # the same streams as an rtlsdr would deliver them (u8 I/Q, 2 bytes per sample): the handler's
    # (b - 127) / 128 conversion is fused into the front-end kernel, PCIe carries a quarter of the bytes
    raw = None
    if not args.no_e2e:
        x8 = torch.view_as_real(x).mul(128.0).add_(127.0).round_().clamp_(0, 255).to(torch.uint8)   # [S, n, 2]
        for _ in range(2):
            proc.process_raw_device(x8.data_ptr(), "u8", 128, n, n, d_audio.data_ptr(), d_audio.stride(0),
                                    d_rds.data_ptr(), d_rds.stride(0))
        ms8 = timed(lambda: proc.process_raw_device(x8.data_ptr(), "u8", 128, n, n, d_audio.data_ptr(),
                                                    d_audio.stride(0), d_rds.data_ptr(), d_rds.stride(0)),
                    args.steps)
        h8 = torch.empty((S, n, 2), dtype=torch.uint8, pin_memory=True)
        h8.copy_(x8)
        del x8

        def step_host_u8():
            rc = proc.L.sdrjfm_process_raw(proc.h, h8.data_ptr(), 1, 128, n, n, ha.data_ptr(), ha.stride(0),
                                           C.byref(na), hr.data_ptr(), hr.stride(0), C.byref(nr), None)
            assert rc == 0, proc.L.sdrjfm_last_error(proc.h)

        step_host_u8()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host_u8()
        barrier()
        dt8 = time.perf_counter() - t0
        dts8 = [dt8]
        if world > 1:
            t = torch.tensor([dt8], device=dev, dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            dts8 = [float(v.item()) for v in allt]
            dt8 = max(dts8)
        raw = {"format": "u8 I/Q as delivered by rtlsdr (2 B/sample), converted in the front-end kernel",
               "value": world * S * n * args.steps / (ms8 * 1e-3) / 1e6,
               "e2e": {"value": world * S * n * e2e_steps / dt8 / 1e6, "unit": "MS/s",
                       "h2d_bytes_per_step": S * n * 2,
                       "d2h_bytes_per_step": S * (na.value + nr.value) * 8, "steps": e2e_steps,
                       "h2d_GBps_per_rank": [S * n * 2 * e2e_steps / d / 1e9 for d in dts8]}}

    def chain_rate(streams, samples, stream0, steps, front_end_mode=0):
        """device-resident rate of a fresh handle with `streams` streams x `samples` IQ samples per step"""
        q = pkg.FmProcessorB200(n_streams=streams, max_samples_per_call=samples, device=local, keep_taps=False,
                                front_end_mode=front_end_mode)
        q.configure(**settings)
        xq = gen_batch_gpu(torch, dev, streams, samples, stream0=stream0)
        da = torch.empty((streams, samples // 48 + 16), dtype=torch.complex64, device=dev)
        dr = torch.empty((streams, samples // 96 + 16), dtype=torch.complex64, device=dev)
        stq = torch.cuda.ExternalStream(q.cuda_stream, device=dev)

        def f():
            q.process_device(xq.data_ptr(), samples, xq.stride(0), da.data_ptr(), da.stride(0), dr.data_ptr(), dr.stride(0))
        for _ in range(3):
            f()
        q.sync()
        msq = timed(f, steps, stq)
        per = [v / steps for v in timed.per_rank]
        q.close()
        del xq, da, dr
        return msq / steps, per

    # BASELINE config 5 as written: `--streams` streams IN TOTAL, sharded over the ranks (strong scaling).  Every
    # rank runs its contiguous share (sharding.stream_range) of the same 256 x n batch the N=1 line processes.
    strong = None
    if not args.no_sweep:
        sh = importlib.import_module("sdrjfm_b200.sharding")
        if world == 1:
            strong = {"streams_total": S, "streams_per_gpu": [S], "ms_per_step": ms / args.steps,
                      "value": value, "per_rank_ms": per_rank_ms, "note": "N=1: the weak-scaling line itself"}
        else:
            lo, hi = sh.stream_range(S, world, rank)
            ms_s, per = chain_rate(hi - lo, n, lo, max(5, args.steps // 2))
            strong = {"streams_total": S, "streams_per_gpu": [sh.stream_range(S, world, r)[1] - sh.stream_range(S, world, r)[0]
                                                               for r in range(world)],
                      "samples_per_stream": n, "ms_per_step": ms_s, "value": S * n / (ms_s * 1e-3) / 1e6, "unit": "MS/s",
                      "per_rank_ms": per, "scaling": "strong"}

    # the chain against the number of streams on ONE GPU, at constant bytes per step (streams x seconds fixed):
    # the per-stream recurrences (pilot PLL, PSS) bound the rate when few streams share the SMs
    sweep_streams = None
    if not args.no_sweep and world == 1:
        sweep_streams = []
        for s_ in (32, 64, 128, 256, 512, 1024):
            if s_ == S:
                sweep_streams.append({"streams": s_, "samples_per_stream": n, "ms_per_step": ms / args.steps, "GS_per_s": value / 1e3})
                continue
            n_ = (S * n // s_) // (12 * 512) * (12 * 512)
            ms_, _ = chain_rate(s_, n_, 0, 5)
            sweep_streams.append({"streams": s_, "samples_per_stream": n_, "ms_per_step": ms_,
                                  "GS_per_s": s_ * n_ / (ms_ * 1e-3) / 1e9})

    # the reference-order front end (front_end_mode 2: fm-rate samples bit-identical to the reference's)
    exact_fe = None
    if not args.no_sweep and world == 1:
        ms_x, _ = chain_rate(S, n, 0, 3, front_end_mode=2)
        exact_fe = {"front_end_mode": 2, "ms_per_step": ms_x, "value": S * n / (ms_x * 1e-3) / 1e6, "unit": "MS/s",
                    "note": "whole chain with the per-sample float32 DC recurrence and the tap-by-tap decimators "
                            "(frontend_exact.cuh); latency-bound by the DC walker"}

    # the GUI's real-time cadence: ONE stream, 16384-sample pulls (7.1 ms of signal,
    # fm-processor.cpp:374) through the host-buffer call, as the Qt adapter makes them
    gui = None
    if not args.no_e2e and rank == 0:
        sig = importlib.import_module("sdrjfm_b200.signals")
        g = pkg.FmProcessorB200(n_streams=1, max_samples_per_call=16384, device=local, keep_taps=False)
        g.configure(**settings)
        xs = sig.batch_stream(0, 16384 * 64)
        ga = np.zeros((1, 16384 // 48 + 2), np.complex64)
        gr = np.zeros((1, 16384 // 96 + 2), np.complex64)
        import ctypes as C2
        a1, r1 = C2.c_int64(0), C2.c_int64(0)
        lat = []
        for i in range(64 * 4):
            blk = xs[(i % 64) * 16384:(i % 64 + 1) * 16384]
            t0 = time.perf_counter()
            rc = g.L.sdrjfm_process(g.h, blk.ctypes.data, 16384, 16384, ga.ctypes.data, ga.shape[1], C2.byref(a1),
                                    gr.ctypes.data, gr.shape[1], C2.byref(r1), None)
            lat.append(time.perf_counter() - t0)
            assert rc == 0
        lat = np.array(lat[32:]) * 1e3
        gui = {"call": "sdrjfm_process, 1 stream x 16384 IQ samples (7.11 ms of signal) from pageable host memory, "
                       "audio + RDS baseband back", "calls": int(len(lat)), "ms_per_call_median": float(np.median(lat)),
               "ms_per_call_p99": float(np.percentile(lat, 99)), "ms_per_call_max": float(lat.max()),
               "x_realtime": float(16384 / 2304000 * 1e3 / np.median(lat))}
        g.close()

    sweep = None
    if args.front_end_sweep and rank == 0:
        sweep = front_end_sweep(pkg, torch, dev, local, S, peak)
    if rank == 0:
        clk.__exit__()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_both(settings, args.cpu_seconds)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {"metric": "IQ MS/s through full FM demod chain", "value": value, "unit": "MS/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "x_realtime_2p304MSps": value / 2.304, "per_rank_ms": per_rank_ms,
                "x_realtime_note": "aggregate over all streams of the batch; ONE stream runs at x_realtime of gui_cadence "
                                   "(the per-stream recurrences are sequential in time)",
                "strong_scaling": strong, "streams_sweep": sweep_streams, "front_end_exact": exact_fe,
                "audio48_parity": audio48_parity(), "roofline": roofline,
                "cpu_baseline": cpu, "e2e": e2e, "device_format_u8": raw, "front_end_sweep": sweep, "gui_cadence": gui,
                "gpu_launches": int(launches),
                "clocks": clk.summary()}
        emit(line)
    proc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
