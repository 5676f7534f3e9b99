"""TEST INFRASTRUCTURE — ctypes loader for the two CPU checkers (oracle/chain_api.h).

`load("ref")` -> oracle/_ref/libsdrjfm_ref.so   (reference classes + ref_harness.cpp)
`load("orc")` -> oracle/_build/libsdrjfm_oracle.so (our restatement, fm_oracle.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing under the product package imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class ChainCfg(C.Structure):
    _fields_ = [
        ("input_rate", C.c_int32), ("fm_rate", C.c_int32), ("fm_mode", C.c_int32),
        ("decoder", C.c_int32), ("sound_sel", C.c_int32), ("rds_on", C.c_int32),
        ("auto_mono", C.c_int32), ("pss_on", C.c_int32), ("dc_remove", C.c_int32),
        ("input_filter_hz", C.c_int32), ("lf_cutoff_hz", C.c_int32), ("lo_hz", C.c_int32),
        ("lgain", C.c_float), ("rgain", C.c_float), ("deemph_us", C.c_int32),
        ("volume_db", C.c_float), ("panorama", C.c_int32), ("balance", C.c_int32),
        ("squelch_mode", C.c_int32), ("squelch_value", C.c_int32),
    ]


class ChainTaps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "fm_z", "demod", "pilot_phase", "locked", "pss_delay", "lr", "audio192",
        "rds_cplx", "rds24")]


class ChainMeta(C.Structure):
    _fields_ = [
        ("dc_rf_re", C.c_float), ("dc_rf_im", C.c_float), ("dc_if", C.c_float),
        ("carrier_ampl", C.c_float), ("pss_phase_shift", C.c_float),
        ("pss_mean_error", C.c_float), ("pss_minimized", C.c_int32),
        ("pilot_lock_strength", C.c_float), ("pilot_locked", C.c_int32),
        ("squelch_active", C.c_int32),
    ]


DEFAULTS = dict(input_rate=2304000, fm_rate=192000, fm_mode=0, decoder=3, sound_sel=0,
                rds_on=0, auto_mono=1, pss_on=1, dc_remove=1, input_filter_hz=0,
                lf_cutoff_hz=0, lo_hz=0, lgain=1.0, rgain=1.0, deemph_us=50,
                volume_db=-6.0, panorama=100, balance=0, squelch_mode=0, squelch_value=0)

DUMP = dict(fmband1=0, fmband2=1, rdsdecim=2, input_filter_freq=3, rds_bp_freq=4,
            pss_lp_freq=5, audio_lp_freq=6, sincos=7, atan=8, consts=9, squelch_iir=10)

_PATHS = {"ref": os.path.join(HERE, "_ref", "libsdrjfm_ref.so"),
          "orc": os.path.join(HERE, "_build", "libsdrjfm_oracle.so")}


def build(which=("oracle", "ref")):
    """make the checkers (the `ref` target is a no-op when /root/reference is absent)."""
    for tgt in which:
        subprocess.run(["make", "-s", "-C", HERE, tgt], check=True)


def available(prefix):
    return os.path.exists(_PATHS[prefix])


def make_cfg(**kw):
    d = dict(DEFAULTS)
    for k, v in kw.items():
        if k not in d:
            raise KeyError(k)
        d[k] = v
    return ChainCfg(**d)


class Chain:
    """One fmProcessor-equivalent instance of a CPU checker (stateful, streaming)."""

    TAPS = ("fm_z", "demod", "pilot_phase", "locked", "pss_delay", "lr", "audio192",
            "rds_cplx", "rds24")

    def __init__(self, prefix, **cfg):
        self.prefix = prefix
        self.lib = C.CDLL(_PATHS[prefix])
        f = lambda n: getattr(self.lib, f"{prefix}_{n}")
        self._create, self._destroy = f("create"), f("destroy")
        self._process, self._meta, self._dump = f("process"), f("get_meta"), f("dump_taps")
        self._process_demod = f("process_demod")
        self._process_demod.restype = C.c_int64
        self._process_demod.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(ChainTaps),
                                        C.POINTER(C.c_int64)]
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.POINTER(ChainCfg)]
        self._destroy.argtypes = [C.c_void_p]
        self._process.restype = C.c_int64
        self._process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(ChainTaps),
                                  C.POINTER(C.c_int64)]
        self._meta.argtypes = [C.c_void_p, C.POINTER(ChainMeta)]
        self._dump.restype = C.c_int32
        self._dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32]
        self.cfg = make_cfg(**cfg)
        self.h = self._create(C.byref(self.cfg))

    def __del__(self):
        if getattr(self, "h", None):
            self._destroy(self.h)
            self.h = None

    def update(self, actions=0, **cfg):
        """run-time setters between process calls (ref_ only): new values for fm_mode, decoder, sound_sel,
        panorama, balance, volume_db, deemph_us, auto_mono, pss_on, lgain, rgain, lo_hz, squelch_*;
        actions: 1 restartPssAnalyzer, 2 triggerFrequencyChange, 4 setDCRemove (dc_remove=...)."""
        assert self.prefix == "ref"
        for k, v in cfg.items():
            setattr(self.cfg, k, v)
        self.lib.ref_update.restype = None
        self.lib.ref_update.argtypes = [C.c_void_p, C.POINTER(ChainCfg), C.c_int32]
        self.lib.ref_update(self.h, C.byref(self.cfg), actions)

    def process_demod(self, demod, taps=TAPS):
        """enter the chain AFTER the discriminator with given float32 demod values."""
        demod = np.ascontiguousarray(demod, dtype=np.float32)
        return self._run(self._process_demod, demod, demod.shape[0], demod.shape[0] + 2, taps)

    def process(self, iq, taps=TAPS):
        """iq: complex64[n]. Returns dict tap -> ndarray (fm-rate length, rds24 shorter)."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        n = iq.shape[0]
        irate = self.cfg.input_rate // 6                      # fm-processor.cpp:36,68-75
        dec = max(1, (self.cfg.input_rate // irate) * (irate // self.cfg.fm_rate)) if self.cfg.input_rate > self.cfg.fm_rate else 1
        return self._run(self._process, iq, n, n // min(dec, 12) + 2, taps)

    def process_fm(self, z, taps=TAPS):
        """for a chain created with input_rate == fm_rate (the reference's decimator bypass,
        fm-processor.cpp:471): z = complex64 samples already at the fm rate."""
        assert self.cfg.input_rate == self.cfg.fm_rate
        z = np.ascontiguousarray(z, dtype=np.complex64)
        return self._run(self._process, z, z.shape[0], z.shape[0] + 2, taps)

    def _run(self, fn, iq, n, cap, taps):
        bufs = {}
        t = ChainTaps()
        for name in taps:
            if name == "locked":
                a = np.zeros(cap, np.uint8)
            elif name in ("fm_z", "lr", "audio192", "rds_cplx", "rds24"):
                a = np.zeros(cap, np.complex64)
            else:
                a = np.zeros(cap, np.float32)
            bufs[name] = a
            setattr(t, name, a.ctypes.data)
        nr = C.c_int64(0)
        nfm = fn(self.h, iq.ctypes.data, n, C.byref(t), C.byref(nr))
        out = {}
        for name, a in bufs.items():
            out[name] = a[:nr.value] if name == "rds24" else a[:nfm]
        out["n_fm"], out["n_rds24"] = nfm, nr.value
        return out

    def meta(self):
        m = ChainMeta()
        self._meta(self.h, C.byref(m))
        return {k: getattr(m, k) for k, _ in ChainMeta._fields_}

    def dump(self, which, cap=200000):
        a = np.zeros(cap, np.complex64)
        n = self._dump(self.h, DUMP[which], a.ctypes.data, cap)
        return None if n < 0 else a[:n].copy()


class Rds1:
    """RDS symbol stage, mode RDS_1: the reference's Costas loop + rdsDecoder_1 (ref_ only)."""

    def __init__(self, rate=24000):
        self.lib = C.CDLL(_PATHS["ref"])
        self.lib.ref_rds1_create.restype = C.c_void_p
        self.lib.ref_rds1_create.argtypes = [C.c_int32]
        self.lib.ref_rds1_destroy.argtypes = [C.c_void_p]
        self.lib.ref_rds1_process.restype = C.c_int64
        self.lib.ref_rds1_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        self.lib.ref_rds1_dump.restype = C.c_int32
        self.lib.ref_rds1_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32]
        self.h = self.lib.ref_rds1_create(rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_rds1_destroy(self.h)
            self.h = None

    def process(self, rds24):
        x = np.ascontiguousarray(rds24, dtype=np.complex64)
        bits = np.zeros(len(x) // 8 + 16, np.uint8)
        n = self.lib.ref_rds1_process(self.h, x.ctypes.data, len(x), bits.ctypes.data, len(bits))
        return bits[:n].copy()

    def dump(self, which):
        a = np.zeros(64, np.float32)
        n = self.lib.ref_rds1_dump(self.h, which, a.ctypes.data, 64)
        return a[:n].copy()


def ref_rds1_mag(rds24, rate=24000):
    """magCplx = 4 x Costas output per 24 kHz sample (mode RDS_1, fresh loop; ref_ only)."""
    lib = C.CDLL(_PATHS["ref"])
    lib.ref_rds1_mag.restype = None
    lib.ref_rds1_mag.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    x = np.ascontiguousarray(rds24, dtype=np.complex64)
    out = np.zeros(len(x), np.complex64)
    lib.ref_rds1_mag(rate, x.ctypes.data, len(x), out.ctypes.data)
    return out


def ref_lf_spectrum(v, spectrum_size=2048, display_size=512, average_count=5, zoom=1, show_full=False):
    """ls_scope's displayBuffer after every block of spectrum_size samples of the LF scope stream v
    (restated around the reference's Fft_transform; ref_ only). Returns float64 [blocks, display_size]."""
    lib = C.CDLL(_PATHS["ref"])
    lib.ref_lf_spectrum.restype = C.c_int64
    lib.ref_lf_spectrum.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p]
    x = np.ascontiguousarray(v, dtype=np.complex64)
    out = np.zeros((len(x) // spectrum_size + 1, display_size), np.float64)
    n = lib.ref_lf_spectrum(x.ctypes.data, len(x), spectrum_size, display_size, average_count, zoom,
                            int(show_full), out.ctypes.data)
    return out[:n].copy()


def ref_scan_blocks(fm_z):
    """(signal dB, noise dB) per 1024-sample block, computed with the reference's FFT (ref_ only)."""
    lib = C.CDLL(_PATHS["ref"])
    lib.ref_scan_blocks.restype = C.c_int64
    lib.ref_scan_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    z = np.ascontiguousarray(fm_z, dtype=np.complex64)
    out = np.zeros((len(z) // 1024 + 1, 2), np.float32)
    n = lib.ref_scan_blocks(z.ctypes.data, len(z), out.ctypes.data)
    return out[:n].copy()


class RefPost:
    """insertTestTone + evaluatePeakLevel (fm-processor.cpp:772-823) as restated in ref_harness.cpp (ref_ only)."""

    def __init__(self, working_rate=48000):
        self.lib = C.CDLL(_PATHS["ref"])
        self.lib.ref_post_create.restype = C.c_void_p
        self.lib.ref_post_create.argtypes = [C.c_int32]
        self.lib.ref_post_destroy.argtypes = [C.c_void_p]
        self.lib.ref_post_set.restype = None
        self.lib.ref_post_set.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        self.lib.ref_post_process.restype = C.c_int64
        self.lib.ref_post_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        self.h = self.lib.ref_post_create(working_rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_post_destroy(self.h)
            self.h = None

    def set(self, tone_on, delay_steps=-1):
        self.lib.ref_post_set(self.h, int(tone_on), int(delay_steps))

    def process(self, pcm):
        """pcm complex64 [n] -> (pcm behind the test tone, [(left dB, right dB)] showPeakLevel read-outs)."""
        x = np.ascontiguousarray(pcm, dtype=np.complex64)
        out = np.zeros_like(x)
        peaks = np.zeros((len(x) // 900 + 4, 2), np.float32)
        n = self.lib.ref_post_process(self.h, x.ctypes.data, len(x), out.ctypes.data, peaks.ctypes.data, len(peaks))
        return out, peaks[:n].copy()


class Rds2:
    """RDS symbol stage, mode RDS_2: the reference's rdsDecoder_2 (matched filter, AGC, M&M timing, Costas) (ref_ only)."""

    def __init__(self, rate=24000):
        self.lib = C.CDLL(_PATHS["ref"])
        self.lib.ref_rds2_create.restype = C.c_void_p
        self.lib.ref_rds2_create.argtypes = [C.c_int32]
        self.lib.ref_rds2_destroy.argtypes = [C.c_void_p]
        self.lib.ref_rds2_process.restype = C.c_int64
        self.lib.ref_rds2_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        self.lib.ref_rds2_dump.restype = C.c_int32
        self.lib.ref_rds2_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        self.h = self.lib.ref_rds2_create(rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_rds2_destroy(self.h)
            self.h = None

    def process(self, rds24):
        x = np.ascontiguousarray(rds24, dtype=np.complex64)
        bits = np.zeros(len(x) // 8 + 16, np.uint8)
        n = self.lib.ref_rds2_process(self.h, x.ctypes.data, len(x), bits.ctypes.data, len(bits))
        return bits[:n].copy()

    def matched_filter(self):
        a = np.zeros(64, np.float32)
        n = self.lib.ref_rds2_dump(self.h, a.ctypes.data, 64)
        return a[:n].copy()


class Rds3:
    """RDS symbol stage, mode RDS_3: the reference's Costas loop, rdsDecoder_3, rdsBlockSynchronizer and RDSGroup (ref_ only)."""

    def __init__(self, rate=24000):
        self.lib = C.CDLL(_PATHS["ref"])
        self.lib.ref_rds3_create.restype = C.c_void_p
        self.lib.ref_rds3_create.argtypes = [C.c_int32]
        self.lib.ref_rds3_destroy.argtypes = [C.c_void_p]
        self.lib.ref_rds3_process.restype = C.c_int64
        self.lib.ref_rds3_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.h = self.lib.ref_rds3_create(rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_rds3_destroy(self.h)
            self.h = None

    def process(self, rds24):
        """-> (bits, groups [n, 4], bit-clock re-synchronisations)"""
        x = np.ascontiguousarray(rds24, dtype=np.complex64)
        bits = np.zeros(len(x) // 8 + 16, np.uint8)
        groups = np.zeros((len(x) // 2000 + 8, 4), np.uint16)
        ng, nrs = C.c_int64(0), C.c_int32(0)
        n = self.lib.ref_rds3_process(self.h, x.ctypes.data, len(x), bits.ctypes.data, len(bits),
                                      groups.ctypes.data, groups.shape[0], C.byref(ng), C.byref(nrs))
        return bits[:n].copy(), groups[:ng.value].copy(), nrs.value


def ref_blocksync_groups(bits):
    """the reference's rdsBlockSynchronizer fed with a bit stream (as rdsDecoder::processBit does) -> completed groups"""
    lib = C.CDLL(_PATHS["ref"])
    lib.ref_blocksync_groups.restype = C.c_int64
    lib.ref_blocksync_groups.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    g = np.zeros((len(b) // 104 + 4, 4), np.uint16)
    n = lib.ref_blocksync_groups(b.ctypes.data, len(b), g.ctypes.data, g.shape[0])
    return g[:n].copy()


def ref_hf_spectrum(x, display_size=1024, sample_rate=2304000, repeat_rate=10):
    """hs_scope's displayBuffer after every completed segment of the raw IQ x (restated around the reference's
    Fft_transform; ref_ only). Returns float64 [segments, display_size]."""
    lib = C.CDLL(_PATHS["ref"])
    lib.ref_hf_spectrum.restype = C.c_int64
    lib.ref_hf_spectrum.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
    z = np.ascontiguousarray(x, dtype=np.complex64)
    cap = len(z) // (sample_rate // repeat_rate) + 1
    out = np.zeros((cap, display_size), np.float64)
    n = lib.ref_hf_spectrum(z.ctypes.data, len(z), display_size, sample_rate, repeat_rate, out.ctypes.data, cap)
    return out[:n].copy()
