"""debug helper (GPU box): per-tap error statistics of the CUDA path vs the CPU checker."""
import sys, os, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
sig = importlib.import_module("sdrjfm_b200.signals")
from oracle import chainlib

def rms(a): return float(np.sqrt(np.mean(np.abs(a.astype(np.complex128))**2)))

def run(x, label, **cfg):
    n = len(x)
    p = pkg.FmProcessorB200(n_streams=1, max_samples_per_call=n)
    p.configure(**cfg)
    p.process(x)
    which = "ref" if chainlib.available("ref") else "orc"
    ref = chainlib.Chain(which, **cfg).process(x)
    print("==", label, cfg)
    for k in ("fm_z", "demod", "pilot_phase", "locked", "lr", "audio192"):
        g = p.read_tap(k); r = ref[k]
        if len(g) != len(r): print(k, "LEN", len(g), len(r)); continue
        d = g.astype(np.complex128) - r.astype(np.complex128)
        if k == "pilot_phase": d = np.angle(np.exp(1j*d.real))
        i = int(np.argmax(np.abs(d)))
        segs = [rms(d[j:j+len(d)//8]) for j in range(0, len(d) - len(d)//8 + 1, len(d)//8)]
        print(f"{k:12s} rms {rms(d):.3e} max {np.abs(d).max():.3e} at {i} (g={g[i]}, r={r[i]}) ref_rms {rms(r):.3e}")
        print("             seg rms:", " ".join(f"{s:.1e}" for s in segs))
        print("             first diffs:", np.abs(d[:6]))
    print("meta gpu", p.meta()[0]); 
    p.close()

n = 2304000 // 4
run(sig.mono_tone(n), "mono_tone", fm_mode=2, volume_db=0.0)
run(sig.dc_offset(sig.mono_tone(n)), "mono_tone+dc", fm_mode=2, volume_db=0.0)
run(sig.mono_tone(n, snr_db=None), "mono clean nodc", fm_mode=2, volume_db=0.0, dc_remove=0)
