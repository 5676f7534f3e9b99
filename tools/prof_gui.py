"""Where the time of ONE GUI pull goes (1 stream x 16384 IQ samples, fm-processor.cpp:374): the whole host call,
the same with pinned host buffers, the device-resident call + sync, launches per call.  Not a bench value."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
pkg = load_package()


QUICK = os.environ.get('PROF_GUI_QUICK') == '1'


def med(f, n=300, skip=40):
    if QUICK:
        n, skip = 30, 5
    t = []
    for i in range(n):
        t0 = time.perf_counter()
        f(i)
        t.append(time.perf_counter() - t0)
    return float(np.median(np.array(t[skip:])) * 1e3)


def main():
    N = 16384
    sig = importlib.import_module("sdrjfm_b200.signals")
    xs = sig.batch_stream(0, N * 64)
    g = pkg.FmProcessorB200(n_streams=1, max_samples_per_call=N, device=0, keep_taps=False)
    g.configure(fm_mode=0, decoder=3, rds_on=1, auto_mono=1, pss_on=1, dc_remove=1, deemph_us=50, volume_db=-6.0)
    ga = np.zeros((1, N // 48 + 2), np.complex64)
    gr = np.zeros((1, N // 96 + 2), np.complex64)
    a1, r1 = C.c_int64(0), C.c_int64(0)

    def host(i):
        blk = xs[(i % 64) * N:(i % 64 + 1) * N]
        rc = g.L.sdrjfm_process(g.h, blk.ctypes.data, N, N, ga.ctypes.data, ga.shape[1], C.byref(a1),
                                gr.ctypes.data, gr.shape[1], C.byref(r1), None)
        assert rc == 0
    l0 = g.launch_count
    out = {"host_pageable_ms": med(host)}
    out["launches_per_call"] = (g.launch_count - l0) / (30.0 if QUICK else 300.0)
    out["pilot_stats_last_call(passes, worst, fallbacks, windows)"] = g.pilot_stats().reshape(-1)[:4].tolist()
    if QUICK:
        print(out)
        g.close()
        return

    xp = torch.from_numpy(xs.view(np.float32).copy()).pin_memory()
    pa = torch.zeros(ga.shape[1] * 2, dtype=torch.float32).pin_memory()
    pr = torch.zeros(gr.shape[1] * 2, dtype=torch.float32).pin_memory()

    def pinned(i):
        rc = g.L.sdrjfm_process(g.h, xp.data_ptr() + (i % 64) * N * 8, N, N, pa.data_ptr(), ga.shape[1], C.byref(a1),
                                pr.data_ptr(), gr.shape[1], C.byref(r1), None)
        assert rc == 0
    out["host_pinned_ms"] = med(pinned)

    xd = torch.from_numpy(xs.view(np.float32).copy()).cuda()
    da = torch.zeros(ga.shape[1] * 2, dtype=torch.float32, device="cuda")
    dr = torch.zeros(gr.shape[1] * 2, dtype=torch.float32, device="cuda")

    def dev(i):
        g.process_device(xd.data_ptr() + (i % 64) * N * 8, N, N, da.data_ptr(), ga.shape[1], dr.data_ptr(), gr.shape[1])
        g.sync()
    out["device_call_plus_sync_ms"] = med(dev)

    def dev_nosync(i):
        g.process_device(xd.data_ptr() + (i % 64) * N * 8, N, N, da.data_ptr(), ga.shape[1], dr.data_ptr(), gr.shape[1])
    t = med(dev_nosync)
    g.sync()
    out["device_call_enqueue_only_ms"] = t

    # multiple of 12: no pending samples to stage
    N2 = 16380

    def host12(i):
        rc = g.L.sdrjfm_process(g.h, xp.data_ptr() + (i % 64) * N * 8, N2, N2, pa.data_ptr(), ga.shape[1], C.byref(a1),
                                pr.data_ptr(), gr.shape[1], C.byref(r1), None)
        assert rc == 0
    g.sync()
    out["host_pinned_16380_ms"] = med(host12)

    s = torch.cuda.Stream()

    def copies(i):
        with torch.cuda.stream(s):
            xd[:N * 2].copy_(xp[:N * 2], non_blocking=True)
            pa.copy_(da, non_blocking=True)
            pr.copy_(dr, non_blocking=True)
        s.synchronize()
    out["three_copies_and_sync_ms"] = med(copies)
    print(out)
    g.close()


if __name__ == "__main__":
    main()
