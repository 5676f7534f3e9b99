#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the REFERENCE's own classes (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on the seeded signals of sdr-j-fm_b200/signals.py.

The reference ships no test vectors of its own; these fixtures freeze what its DSP
classes produce so that the port (oracle/fm_oracle.cpp) and the CUDA path stay pinned on machines
where /root/reference — and with it oracle/_ref — does not exist.  Run from the repo root in a
container that has /root/reference:      python tests/golden/make_golden.py

Per case the file holds, for every output tap, the SHA-256 of the full float32 stream (bit-exact pin
for the port) and the last TAIL samples themselves (tolerance pin for the CUDA path: 1e-5 RMS).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import chainlib  # noqa: E402

TAIL = 9600          # fm-rate samples kept per tap (50 ms); rds24: TAIL // 8
TAPS = ("fm_z", "demod", "pilot_phase", "locked", "pss_delay", "lr", "audio192", "rds_cplx", "rds24")

# name -> (signal generator, generator kwargs, input samples, chain_cfg)   [BASELINE.json configs 1-5 + AM]
CASES = {
    "config1_mono": ("mono_tone", {}, 2304000 // 4, dict(fm_mode=2, volume_db=0.0)),
    "config2_stereo_pss": ("stereo_pilot", {}, 2304000 * 7 // 10, dict(fm_mode=0, volume_db=0.0)),
    "config3_input_filter": ("adjacent_interferer", {}, 2304000 * 7 // 10,
                             dict(fm_mode=0, input_filter_hz=165000, volume_db=0.0)),
    "config4_rate_6M": ("stereo_pilot", dict(fs=192000 * 30), 6000000 * 7 // 10,
                        dict(input_rate=6000000, fm_mode=0, volume_db=0.0)),
    "config5_batch_rds": ("batch_stream", dict(s=7), 2304000 * 7 // 10, dict(fm_mode=0, rds_on=1, volume_db=-6.0)),
    "am_decoder": ("am_tone", {}, 2304000 // 4, dict(decoder=1, fm_mode=2, volume_db=0.0)),
}


def make_input(sig, name):
    gen, kw, n, _ = CASES[name]
    kw = dict(kw)
    if gen == "batch_stream":
        return sig.batch_stream(kw.pop("s"), n, **kw)
    return getattr(sig, gen)(n, **kw)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import importlib
    load_package()
    sig = importlib.import_module("sdrjfm_b200.signals")
    chainlib.build(("oracle", "ref"))
    assert chainlib.available("ref"), "oracle/_ref is needed: run where /root/reference exists"
    out, meta = {}, {}
    for name, (_, _, n, cfg) in CASES.items():
        x = make_input(sig, name)
        c = chainlib.Chain("ref", **cfg)
        o = c.process(x)
        meta[name] = {"cfg": cfg, "n_in": int(n), "n_fm": int(o["n_fm"]), "n_rds24": int(o["n_rds24"]),
                      "input_sha256": sha(x), "sha256": {}}
        for t in TAPS:
            a = o[t]
            meta[name]["sha256"][t] = sha(a)
            tail = TAIL // 8 if t == "rds24" else TAIL
            out[f"{name}/{t}"] = a[-tail:] if len(a) else a
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
