#!/usr/bin/env python
"""Runs ONLY the decimating front end (one launch over all streams, no lanes) — the command to put under ncu.
   python tools/prof_frontend.py [--format cf32|u8|s8|s16] [--rate 2304000] [--mode 0|1|2] [--streams 256] [--seconds 0.5] [--reps 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SDRJFM_LANES"] = "1"
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--format", default="cf32")
    ap.add_argument("--rate", type=int, default=2304000)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=0.5)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    pkg = bench.load_package()
    dev = torch.device("cuda", 0)
    dec = 5 if a.mode == 1 else pkg.front_end_decimation(a.rate)
    n = int(a.seconds * a.rate) // (dec * 512) * (dec * 512)
    p = pkg.FmProcessorB200(n_streams=a.streams, input_rate=a.rate, max_samples_per_call=n, keep_taps=False, front_end_mode=a.mode)
    p.configure(**bench.chain_settings())
    if a.format == "cf32":
        buf = torch.randn((a.streams, n, 2), device=dev, dtype=torch.float32) * 0.3
    else:
        dt = {"u8": torch.uint8, "s8": torch.int8, "s16": torch.int16}[a.format]
        buf = torch.randint(0, 120, (a.streams, n, 2), device=dev, dtype=dt)
    st = torch.cuda.ExternalStream(p.cuda_stream, device=dev)
    for _ in range(2):
        p.run_frontend_only_raw(buf.data_ptr(), a.format, 2048, n, n)
    p.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(a.reps):
            p.run_frontend_only_raw(buf.data_ptr(), a.format, 2048, n, n)
        e1.record(st)
    torch.cuda.synchronize()
    bps = {"cf32": 8, "s16": 4}.get(a.format, 2)
    ms = e0.elapsed_time(e1) / a.reps
    print(f"{a.format} @{a.rate} mode {a.mode}: {ms:.4f} ms per launch, {a.streams * n / ms / 1e6:.0f} GS/s, "
          f"{a.streams * n * (bps + 8.0 / dec) / ms / 1e6:.0f} GB/s algorithmic; streams {a.streams} samples {n}")
    p.close()


if __name__ == "__main__":
    main()
