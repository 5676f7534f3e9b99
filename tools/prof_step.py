#!/usr/bin/env python
"""Runs a few steps of the benchmark chain and nothing else — the command to put under ncu.
   python tools/prof_step.py [--front-end-mode 0|2] [--streams 256] [--seconds 0.5] [--steps 2] [--lanes N]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--front-end-mode", type=int, default=0)
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=0.5)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--decoder", type=int, default=3)
    a = ap.parse_args()
    if a.lanes:
        os.environ["SDRJFM_LANES"] = str(a.lanes)
    import torch
    pkg = bench.load_package()
    dev = torch.device("cuda", 0)
    n = int(a.seconds * bench.INPUT_RATE) // 12 * 12
    st = bench.chain_settings()
    st["decoder"] = a.decoder
    p = pkg.FmProcessorB200(n_streams=a.streams, max_samples_per_call=n, keep_taps=False, front_end_mode=a.front_end_mode)
    p.configure(**st)
    x = bench.gen_batch_gpu(torch, dev, a.streams, n)
    da = torch.empty((a.streams, n // 48 + 16), dtype=torch.complex64, device=dev)
    dr = torch.empty((a.streams, n // 96 + 16), dtype=torch.complex64, device=dev)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(p.cuda_stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(a.warmup + a.steps):
        if i == a.warmup:
            p.sync()
            with torch.cuda.stream(ext):
                e0.record(ext)
        p.process_device(x.data_ptr(), n, x.stride(0), da.data_ptr(), da.stride(0), dr.data_ptr(), dr.stride(0))
    with torch.cuda.stream(ext):
        e1.record(ext)
    p.sync()
    torch.cuda.synchronize()
    print("ms per step", e0.elapsed_time(e1) / a.steps, "launches", p.launch_count)
    if os.environ.get("SDRJFM_DC_STATS"):
        st_ = p.pilot_stats().reshape(-1)
        print("dc solver (passes, worst) first streams:", st_[:16], "windows per call", (n + 3071) // 3072)
    p.close()


if __name__ == "__main__":
    main()
