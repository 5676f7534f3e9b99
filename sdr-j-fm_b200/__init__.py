"""sdr-j-fm_b200 — host-side mirror (Python, ctypes) of the reference's fmProcessor
interface over the C ABI in include/sdrjfm_b200.h.

The directory name carries a hyphen (it is the repo's package directory, not an
installable distribution); load it with `load_package()` from `__graft_entry__` or
`tests/conftest.py`, which registers it as module `sdrjfm_b200`.

There is no CPU fallback anywhere in this package: if the CUDA library is missing or no
B200 is visible, construction raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libsdrjfm_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "sdrjfm_b200.h")

OK, ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

TAP = dict(fm_z=0, demod=1, pilot_phase=2, locked=3, pss_delay=4, lr=5, audio192=6,
           rds_cplx=7, rds24=8, fe_u=9, fe_s=10)
_TAP_DTYPE = dict(fm_z=np.complex64, demod=np.float32, pilot_phase=np.float32,
                  locked=np.uint8, pss_delay=np.float32, lr=np.complex64,
                  audio192=np.complex64, rds_cplx=np.complex64, rds24=np.complex64,
                  fe_u=np.complex64, fe_s=np.complex64)


# device sample formats (enum sdrjfm_iq_format): dtype of one I or Q component
IQ_FORMAT = dict(cf32=0, u8=1, s8=2, s16=3, airspy=4)
# enum class ELfPlot, includes/fm/fm-processor.h:84-86
LF_PLOT = dict(OFF=0, IF_FILTERED=1, DEMODULATOR=2, AF_SUM=3, AF_DIFF=4, AF_MONO_FILTERED=5,
               AF_LEFT_FILTERED=6, AF_RIGHT_FILTERED=7, RDS_INPUT=8, RDS_DEMOD=9)
_IQ_DTYPE = {0: np.float32, 1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int16}


def front_end_decimation(input_rate, fm_rate=192000):
    """input samples per fm-rate sample, from the reference's constructor arithmetic
    (fm-processor.cpp:36,68-75): IRate = inputRate/6; stage 1 /6; stage 2 /(IRate/fmRate)."""
    irate = input_rate // 6
    return (input_rate // irate) * (irate // fm_rate)


class SdrjfmError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"sdrjfm status {status}: {msg}")
        self.status = status


class Config(C.Structure):
    _fields_ = [("input_rate", C.c_int32), ("fm_rate", C.c_int32),
                ("working_rate", C.c_int32), ("audio_rate", C.c_int32),
                ("n_streams", C.c_int32), ("device", C.c_int32),
                ("max_samples_per_call", C.c_int64), ("keep_taps", C.c_int32),
                ("front_end_mode", C.c_int32)]


class Meta(C.Structure):
    _fields_ = [("dc_rf_re", C.c_float), ("dc_rf_im", C.c_float), ("dc_rf_db", C.c_float),
                ("dc_if", C.c_float), ("carrier_ampl", C.c_float),
                ("pss_phase_shift_deg", C.c_float), ("pss_phase_change", C.c_float),
                ("pss_state", C.c_int32), ("pilot_lock_strength", C.c_float),
                ("pilot_locked", C.c_int32), ("peak_left_db", C.c_float),
                ("peak_right_db", C.c_float), ("squelch_active", C.c_int32)]


def build(verbose=False):
    """Compile csrc/ for sm_100a into libsdrjfm_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc")], capture_output=True,
                       text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("building libsdrjfm_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library (loud failure if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build() first; "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        L.sdrjfm_create.restype = vp
        L.sdrjfm_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_int)]
        L.sdrjfm_destroy.argtypes = [vp]
        L.sdrjfm_last_error.restype = C.c_char_p
        L.sdrjfm_last_error.argtypes = [vp]
        L.sdrjfm_version.restype = C.c_char_p
        L.sdrjfm_process.argtypes = [vp, vp, i64, i64, vp, i64, C.POINTER(i64), vp, i64,
                                     C.POINTER(i64), vp]
        L.sdrjfm_process_device.argtypes = [vp, vp, i64, i64, vp, i64, C.POINTER(i64), vp, i64,
                                            C.POINTER(i64)]
        L.sdrjfm_process_raw.argtypes = [vp, vp, i32, i32, i64, i64, vp, i64, C.POINTER(i64), vp, i64,
                                         C.POINTER(i64), vp]
        L.sdrjfm_process_raw_device.argtypes = [vp, vp, i32, i32, i64, i64, vp, i64, C.POINTER(i64),
                                                vp, i64, C.POINTER(i64)]
        L.sdrjfm_run_frontend_only_raw.argtypes = [vp, vp, i32, i32, i64, i64]
        L.sdrjfm_sync.argtypes = [vp]
        L.sdrjfm_get_meta.argtypes = [vp, vp]
        L.sdrjfm_read_tap.restype = i64
        L.sdrjfm_read_tap.argtypes = [vp, C.c_int, i32, vp, i64]
        L.sdrjfm_cuda_stream.restype = vp
        L.sdrjfm_cuda_stream.argtypes = [vp]
        L.sdrjfm_run_frontend_only.argtypes = [vp, vp, i64, i64]
        L.sdrjfm_launch_count.restype = i64
        L.sdrjfm_launch_count.argtypes = [vp]
        L.sdrjfm_pilot_stats.argtypes = [vp, vp]
        for name in ("fm_mode", "fm_decoder", "sound_mode", "stereo_panorama", "sound_balance",
                     "deemphasis", "lf_cutoff", "bandwidth", "rds_mode", "local_oscillator",
                     "squelch_mode", "squelch_value", "native_rate", "rds_symbol_stage", "scanning", "auto_mono", "pss_mode",
                     "dc_remove", "test_tone", "disp_delay"):
            getattr(L, f"sdrjfm_set_{name}").argtypes = [vp, i32]
        L.sdrjfm_set_lf_plot_type.restype = i32
        L.sdrjfm_set_lf_plot_type.argtypes = [vp, i32]
        L.sdrjfm_read_lf_plot.restype = i64
        L.sdrjfm_read_lf_plot.argtypes = [vp, i32, vp, i64, vp, vp]
        L.sdrjfm_set_lf_spectrum.restype = i32
        L.sdrjfm_set_lf_spectrum.argtypes = [vp, i32, i32, i32]
        L.sdrjfm_set_lf_plot_zoom.restype = i32
        L.sdrjfm_set_lf_plot_zoom.argtypes = [vp, i32]
        L.sdrjfm_read_lf_spectrum.restype = i64
        L.sdrjfm_read_lf_spectrum.argtypes = [vp, i32, vp, i64, vp]
        L.sdrjfm_set_hf_spectrum.restype = i32
        L.sdrjfm_set_hf_spectrum.argtypes = [vp, i32, i32]
        L.sdrjfm_read_hf_spectrum.restype = i64
        L.sdrjfm_read_hf_spectrum.argtypes = [vp, i32, vp, i64, vp]
        L.sdrjfm_read_scan.restype = i64
        L.sdrjfm_read_scan.argtypes = [vp, i32, vp, i64]
        L.sdrjfm_read_rds_bits.restype = i64
        L.sdrjfm_read_rds_bits.argtypes = [vp, i32, vp, i64]
        L.sdrjfm_read_rds_groups.restype = i64
        L.sdrjfm_read_rds_groups.argtypes = [vp, i32, vp, i64, vp]
        L.sdrjfm_read_peak_levels.restype = i64
        L.sdrjfm_read_peak_levels.argtypes = [vp, i32, vp, i64]
        L.sdrjfm_set_volume_db.argtypes = [vp, f32]
        L.sdrjfm_set_attenuation.argtypes = [vp, f32, f32]
        L.sdrjfm_trigger_frequency_change.argtypes = [vp]
        L.sdrjfm_restart_pss_analyzer.argtypes = [vp]
        L.sdrjfm_design_tables.restype = i64
        L.sdrjfm_design_tables.argtypes = [i32, i32, i32, i32, vp, i64]
        L.sdrjfm_design_aux.restype = i64
        L.sdrjfm_design_aux.argtypes = [i32, i32, i32, vp, i64]
        L.sdrjfm_tables_nbytes.restype = i64
        L.sdrjfm_tables_nbytes.argtypes = [vp]
        L.sdrjfm_tables_export.argtypes = [vp, vp, i64]
        L.sdrjfm_tables_import.argtypes = [vp, vp, i64]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------
# table blob (TableHeader in csrc/tables.hpp)
_HDR_FIELDS = [("magic", "u4"), ("version", "u4"), ("input_rate", "i4"), ("fm_rate", "i4"),
               ("decim1", "i4"), ("decim2", "i4"), ("ntaps1", "i4"), ("ntaps2", "i4"),
               ("ncomp", "i4"), ("input_filter_hz", "i4"), ("audio_lp_hz", "i4"),
               ("reserved0", "i4"), ("payload_floats", "i8"),
               ("off_fmband1", "i8"), ("off_fmband2", "i8"), ("off_rdsdecim", "i8"),
               ("off_comp", "i8"), ("off_comp_consts", "i8"), ("off_atan", "i8"),
               ("off_sincos", "i8"), ("off_arcsine", "i8"), ("off_tw2048", "i8"),
               ("off_tw8192", "i8"), ("off_tw32768", "i8"), ("off_pss_lp", "i8"),
               ("off_rds_bp", "i8"), ("off_audio_lp", "i8"), ("off_input_taps", "i8"),
               ("off_comp_wide", "i8"), ("ncomp_wide", "i4"), ("reserved1", "i4"),
               ("rs_L", "i4"), ("rs_M", "i4"), ("rs_P", "i4"), ("rs_ntapsA", "i4"),
               ("off_rsA", "i8"), ("off_rsB", "i8"), ("off_squelch", "i8"), ("off_rds_sym", "i8")]
_HDR_DTYPE = np.dtype(_HDR_FIELDS)


class Tables:
    """Parsed view of the shared tap/LUT blob (designed once, broadcast to every rank)."""

    def __init__(self, blob):
        self.blob = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
        self.hdr = self.blob[:_HDR_DTYPE.itemsize].view(_HDR_DTYPE)[0]
        assert self.hdr["magic"] == 0x54464A53
        self.payload = self.blob[_HDR_DTYPE.itemsize:].view(np.float32)

    def _c(self, off, n):
        return self.payload[off:off + 2 * n].view(np.complex64)

    def _f(self, off, n):
        return self.payload[off:off + n]

    @property
    def fmband1(self): return self._c(self.hdr["off_fmband1"], self.hdr["ntaps1"])
    @property
    def fmband2(self): return self._c(self.hdr["off_fmband2"], self.hdr["ntaps2"])
    @property
    def rdsdecim(self): return self._c(self.hdr["off_rdsdecim"], 11)
    @property
    def composite(self): return self._f(self.hdr["off_comp"], self.hdr["ncomp"])
    @property
    def consts(self): return self._f(self.hdr["off_comp_consts"], 8)
    @property
    def atan(self): return self._f(self.hdr["off_atan"], 8 * 8193)
    @property
    def sincos(self): return self._c(self.hdr["off_sincos"], self.hdr["fm_rate"])
    @property
    def pss_lp_freq(self): return self._c(self.hdr["off_pss_lp"], 2048)
    @property
    def rds_bp_freq(self): return self._c(self.hdr["off_rds_bp"], 32768)
    @property
    def audio_lp_freq(self):
        return self._c(self.hdr["off_audio_lp"], 8192) if self.hdr["audio_lp_hz"] > 0 else None
    @property
    def resampler(self):
        """(L, M, P, stage-A taps [49], stage-B taps [L, P]) of the rational polyphase resampler, or None."""
        L, M, P = int(self.hdr["rs_L"]), int(self.hdr["rs_M"]), int(self.hdr["rs_P"])
        if L == 0:
            return None
        return (L, M, P, self._f(self.hdr["off_rsA"], self.hdr["rs_ntapsA"]),
                self._f(self.hdr["off_rsB"], L * P).reshape(L, P))

    @property
    def squelch_iir(self):
        """82 floats: high-pass gain, 10 x (A1 A2 B1 B2), low-pass gain, 10 x (A1 A2 B1 B2)."""
        return self._f(self.hdr["off_squelch"], 82)

    @property
    def rds_symbol(self):
        """(match kernel [43], rdsFilter taps [21], sharpFilter gain + 8 x (A1 A2 B1 B2) [33])."""
        f = self._f(self.hdr["off_rds_sym"], 97)
        return f[:43], f[43:64], f[64:97]

    @property
    def input_taps(self):
        return self._f(self.hdr["off_input_taps"], 251) if self.hdr["input_filter_hz"] > 0 else None


def design_tables(input_rate=2304000, fm_rate=192000, input_filter_hz=0, audio_lp_hz=0):
    """Host-only table design (no GPU needed) -> Tables."""
    L = lib()
    n = L.sdrjfm_design_tables(input_rate, fm_rate, input_filter_hz, audio_lp_hz, None, 0)
    if n < 0:
        raise SdrjfmError(n, "design_tables")
    buf = (C.c_uint8 * n)()
    L.sdrjfm_design_tables(input_rate, fm_rate, input_filter_hz, audio_lp_hz, buf, n)
    return Tables(buf)


def design_aux(which, a, b=0, cap=70000):
    """host-only designers outside the blob: 0 RDS_2 matched filter, 1 test-tone burst, 2 second converter."""
    out = np.zeros(cap, np.float32)
    n = lib().sdrjfm_design_aux(which, a, b, out.ctypes.data, cap)
    if n < 0:
        raise SdrjfmError(n, "design_aux")
    return out[:n].copy()


# ------------------------------------------------------------------------------------------
class FmProcessorB200:
    """GPU counterpart of `fmProcessor` (includes/fm/fm-processor.h:103-157) for a batch of
    `n_streams` independent IQ streams.  Method names follow the reference's setters; the
    data path is `process()` = what fmProcessor::run does with the samples it pulled from
    `deviceHandler::getSamples`, returning what it would push to `audioSink::putSample`
    (48 kHz stereo, complex64 re=L im=R) and to `rdsDecoder::doDecode` (24 kHz complex)."""

    FM_MODE = dict(Stereo=0, StereoPano=1, Mono=2)
    DECODER = {"AM": 1, "FM PLL Decoder": 2, "FM Mixed Demod": 3,
               "FM Complex Baseband Delay": 4, "FM Real Baseband Delay": 5,
               "FM Difference Based": 6}   # fm-demodulator.cpp:36-44

    def __init__(self, n_streams=1, input_rate=2304000, fm_rate=192000, working_rate=48000,
                 audio_rate=48000, max_samples_per_call=2304000, device=0, keep_taps=True,
                 front_end_mode=0):
        self.L = lib()
        self.cfg = Config(input_rate, fm_rate, working_rate, audio_rate, n_streams, device,
                          max_samples_per_call, 1 if keep_taps else 0, front_end_mode)
        st = C.c_int(0)
        self.h = self.L.sdrjfm_create(C.byref(self.cfg), C.byref(st))
        if not self.h:
            raise SdrjfmError(st.value, self.L.sdrjfm_last_error(None).decode())
        self.n_streams = n_streams
        self._display = 0
        self.audio_ratio = audio_rate / working_rate
        # input samples per fm-rate sample (a lower bound in the resampler mode: sizes the outputs)
        self.decim = (input_rate // fm_rate) if front_end_mode == 1 else front_end_decimation(input_rate, fm_rate)

    def close(self):
        if getattr(self, "h", None):
            self.L.sdrjfm_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != OK:
            raise SdrjfmError(rc, self.L.sdrjfm_last_error(self.h).decode())

    # --- setters, same names as the reference -------------------------------------------
    def setfmMode(self, m): self._ck(self.L.sdrjfm_set_fm_mode(self.h, self.FM_MODE.get(m, m)))
    def setFMdecoder(self, d): self._ck(self.L.sdrjfm_set_fm_decoder(self.h, self.DECODER.get(d, d)))
    def setSoundMode(self, s): self._ck(self.L.sdrjfm_set_sound_mode(self.h, s))
    def setStereoPanorama(self, p): self._ck(self.L.sdrjfm_set_stereo_panorama(self.h, p))
    def setSoundBalance(self, b): self._ck(self.L.sdrjfm_set_sound_balance(self.h, b))
    def setDeemphasis(self, us): self._ck(self.L.sdrjfm_set_deemphasis(self.h, us))
    def setVolume(self, db): self._ck(self.L.sdrjfm_set_volume_db(self.h, db))
    def setlfcutoff(self, hz): self._ck(self.L.sdrjfm_set_lf_cutoff(self.h, hz))
    def setBandwidth(self, hz): self._ck(self.L.sdrjfm_set_bandwidth(self.h, hz))
    def setAttenuation(self, l, r): self._ck(self.L.sdrjfm_set_attenuation(self.h, l, r))
    def setfmRdsSelector(self, m): self._ck(self.L.sdrjfm_set_rds_mode(self.h, m))
    def set_localOscillator(self, hz): self._ck(self.L.sdrjfm_set_local_oscillator(self.h, hz))
    def set_squelchMode(self, m): self._ck(self.L.sdrjfm_set_squelch_mode(self.h, m))
    def set_squelchValue(self, n): self._ck(self.L.sdrjfm_set_squelch_value(self.h, n))
    def set_nativeRate(self, hz): self._ck(self.L.sdrjfm_set_native_rate(self.h, hz))
    def startScanning(self): self._ck(self.L.sdrjfm_set_scanning(self.h, 1))
    def stopScanning(self): self._ck(self.L.sdrjfm_set_scanning(self.h, 0))

    def read_scan(self, stream=0):
        """[(signal dB, noise dB)] per 1024-sample block completed by the last process call while scanning."""
        a = np.zeros((self.cfg.max_samples_per_call // (self.decim * 1024) + 4, 2), np.float32)
        n = self.L.sdrjfm_read_scan(self.h, stream, a.ctypes.data, a.shape[0])
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy()

    def setlfPlotType(self, m):
        """fmProcessor::setlfPlotType (ELfPlot name or number; None / -1: no scope stream)."""
        t = -1 if m is None else (LF_PLOT[m] if isinstance(m, str) else int(m))
        self._ck(self.L.sdrjfm_set_lf_plot_type(self.h, t))

    def read_lf_plot(self, stream=0):
        """(complex64 samples pushed into spectrumBuffer_lf by the last call, spectrumSampleRate, showFullSpectrum)."""
        a = np.zeros(self.cfg.max_samples_per_call // self.decim + 8, np.complex64)
        rate = C.c_int32(0); full = C.c_int32(0)
        n = self.L.sdrjfm_read_lf_plot(self.h, stream, a.ctypes.data, a.size, C.byref(rate), C.byref(full))
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy(), rate.value, bool(full.value)

    def setlfPlotZoomFactor(self, zoom): self._ck(self.L.sdrjfm_set_lf_plot_zoom(self.h, int(zoom)))

    def set_lf_spectrum(self, spectrum_size=2048, display_size=512, average_count=5):
        """ls_scope (spectrumSize, displaySize, averageCount) on the GPU; spectrum_size 0 switches it off."""
        self._ck(self.L.sdrjfm_set_lf_spectrum(self.h, spectrum_size, display_size, average_count))
        self._display = display_size

    def read_lf_spectrum(self, stream=0):
        """(displayBuffer float64 [display_size], blocks completed by the last call)."""
        if not self._display:
            raise SdrjfmError(ERR_ARG, "the LF spectrum is off: call set_lf_spectrum first")
        a = np.zeros(self._display, np.float64)
        nb = C.c_int32(0)
        n = self.L.sdrjfm_read_lf_spectrum(self.h, stream, a.ctypes.data, a.size, C.byref(nb))
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy(), nb.value

    def set_hf_spectrum(self, display_size=1024, repeat_rate=10):
        """hs_scope (displaySize, ..., inputRate, repeatRate) on the GPU; display_size 0 switches it off."""
        self._ck(self.L.sdrjfm_set_hf_spectrum(self.h, display_size, repeat_rate))
        self._hf_display = display_size

    def read_hf_spectrum(self, stream=0):
        """(displayBuffer float64 [display_size], segments completed by the last call)."""
        if not getattr(self, "_hf_display", 0):
            raise SdrjfmError(ERR_ARG, "the HF spectrum is off: call set_hf_spectrum first")
        a = np.zeros(self._hf_display, np.float64)
        nb = C.c_int32(0)
        n = self.L.sdrjfm_read_hf_spectrum(self.h, stream, a.ctypes.data, a.size, C.byref(nb))
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy(), nb.value

    def setRdsSymbolStage(self, on): self._ck(self.L.sdrjfm_set_rds_symbol_stage(self.h, int(on)))

    def read_rds_bits(self, stream=0):
        """bits (uint8 0/1) the GPU symbol stage decoded for `stream` in the last process call."""
        a = np.zeros(self.cfg.max_samples_per_call // (8 * self.decim * 16) + 64, np.uint8)
        n = self.L.sdrjfm_read_rds_bits(self.h, stream, a.ctypes.data, a.size)
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy()
    def read_rds_groups(self, stream=0):
        """mode RDS_3: (groups uint16 [n, 4] the device-side block synchroniser completed in the last call,
        dict(synchronized, bitclk_resyncs, sync_errors, crc_errors))."""
        a = np.zeros((self.cfg.max_samples_per_call // (8 * self.decim * 2000) + 8, 4), np.uint16)
        st = (C.c_int32 * 4)()
        n = self.L.sdrjfm_read_rds_groups(self.h, stream, a.ctypes.data, a.shape[0], st)
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy(), dict(synchronized=st[0], bitclk_resyncs=st[1], sync_errors=st[2], crc_errors=st[3])

    def setTestTone(self, on): self._ck(self.L.sdrjfm_set_test_tone(self.h, int(on)))
    def setDispDelay(self, steps): self._ck(self.L.sdrjfm_set_disp_delay(self.h, int(steps)))

    def read_peak_levels(self, stream=0):
        """[(left dB, right dB)] per showPeakLevel read-out of the last process call (one per 961 PCM samples)."""
        a = np.zeros((1024, 2), np.float32)
        n = self.L.sdrjfm_read_peak_levels(self.h, stream, a.ctypes.data, a.shape[0])
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy()

    def setAutoMonoMode(self, on): self._ck(self.L.sdrjfm_set_auto_mono(self.h, int(on)))
    def setPSSMode(self, on): self._ck(self.L.sdrjfm_set_pss_mode(self.h, int(on)))
    def setDCRemove(self, on): self._ck(self.L.sdrjfm_set_dc_remove(self.h, int(on)))
    def triggerFrequencyChange(self): self._ck(self.L.sdrjfm_trigger_frequency_change(self.h))
    def restartPssAnalyzer(self): self._ck(self.L.sdrjfm_restart_pss_analyzer(self.h))

    def configure(self, **kw):
        """Apply the chain settings used by the CPU checkers' chain_cfg (oracle/chain_api.h)."""
        m = dict(fm_mode=self.setfmMode, decoder=self.setFMdecoder, sound_sel=self.setSoundMode,
                 rds_on=self.setfmRdsSelector, auto_mono=self.setAutoMonoMode,
                 pss_on=self.setPSSMode, dc_remove=self.setDCRemove,
                 input_filter_hz=self.setBandwidth, lf_cutoff_hz=self.setlfcutoff,
                 lo_hz=self.set_localOscillator, deemph_us=self.setDeemphasis,
                 squelch_mode=self.set_squelchMode, squelch_value=self.set_squelchValue,
                 volume_db=self.setVolume, panorama=self.setStereoPanorama,
                 balance=self.setSoundBalance)
        if "lgain" in kw or "rgain" in kw:
            self.setAttenuation(kw.pop("lgain", 1.0), kw.pop("rgain", 1.0))
        for k, v in kw.items():
            m[k](v)

    # --- data path -----------------------------------------------------------------------
    def process(self, iq, want_meta=False):
        """iq: complex64 [n_streams, n] (or [n] for one stream), HOST memory.
        Returns (audio48k complex64 [n_streams, n_audio], rds24 complex64 [n_streams, n_rds])."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq[None, :]
        assert iq.shape[0] == self.n_streams
        n = iq.shape[1]
        audio = np.zeros((self.n_streams, int((n // (4 * self.decim) + 2) * self.audio_ratio) + 2), np.complex64)
        rds = np.zeros((self.n_streams, n // (8 * self.decim) + 2), np.complex64)
        na, nr = C.c_int64(0), C.c_int64(0)
        meta = (Meta * self.n_streams)() if want_meta else None
        self._ck(self.L.sdrjfm_process(self.h, iq.ctypes.data, n, iq.shape[1],
                                       audio.ctypes.data, audio.shape[1], C.byref(na),
                                       rds.ctypes.data, rds.shape[1], C.byref(nr),
                                       C.byref(meta) if meta is not None else None))
        out = (audio[:, :na.value], rds[:, :nr.value])
        if want_meta:
            return out + ([{k: getattr(m, k) for k, _ in Meta._fields_} for m in meta],)
        return out

    def process_raw(self, iq, fmt, denominator=2048):
        """iq: [n_streams, n, 2] (or [n, 2]) array of uint8 / int8 / int16 components in HOST memory,
        the bytes a device delivers; fmt in IQ_FORMAT.  Same results as process() on the floats the
        reference's handler would have made of them."""
        f = IQ_FORMAT.get(fmt, fmt)
        iq = np.ascontiguousarray(iq, dtype=_IQ_DTYPE[f])
        if iq.ndim == 2:
            iq = iq[None]
        assert iq.shape[0] == self.n_streams and iq.shape[2] == 2
        n = iq.shape[1]
        audio = np.zeros((self.n_streams, n // (4 * self.decim) + 2), np.complex64)
        rds = np.zeros((self.n_streams, n // (8 * self.decim) + 2), np.complex64)
        na, nr = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.sdrjfm_process_raw(self.h, iq.ctypes.data, f, denominator, n, n,
                                           audio.ctypes.data, audio.shape[1], C.byref(na),
                                           rds.ctypes.data, rds.shape[1], C.byref(nr), None))
        return audio[:, :na.value], rds[:, :nr.value]

    def process_raw_device(self, d_iq_ptr, fmt, denominator, n_in, in_pitch, d_audio_ptr=None,
                           audio_pitch=0, d_rds_ptr=None, rds_pitch=0):
        na, nr = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.sdrjfm_process_raw_device(self.h, d_iq_ptr, IQ_FORMAT.get(fmt, fmt), denominator,
                                                  n_in, in_pitch, d_audio_ptr, audio_pitch, C.byref(na),
                                                  d_rds_ptr, rds_pitch, C.byref(nr)))
        return na.value, nr.value

    def run_frontend_only_raw(self, d_iq_ptr, fmt, denominator, n_in, in_pitch):
        self._ck(self.L.sdrjfm_run_frontend_only_raw(self.h, d_iq_ptr, IQ_FORMAT.get(fmt, fmt), denominator,
                                                     n_in, in_pitch))

    def process_device(self, d_iq_ptr, n_in, in_pitch, d_audio_ptr=None, audio_pitch=0,
                       d_rds_ptr=None, rds_pitch=0):
        """Device-pointer form (asynchronous on the handle's stream). Returns (n_audio, n_rds)."""
        na, nr = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.sdrjfm_process_device(self.h, d_iq_ptr, n_in, in_pitch, d_audio_ptr,
                                              audio_pitch, C.byref(na), d_rds_ptr, rds_pitch,
                                              C.byref(nr)))
        return na.value, nr.value

    def run_frontend_only(self, d_iq_ptr, n_in, in_pitch):
        self._ck(self.L.sdrjfm_run_frontend_only(self.h, d_iq_ptr, n_in, in_pitch))

    def sync(self): self._ck(self.L.sdrjfm_sync(self.h))

    @property
    def cuda_stream(self): return self.L.sdrjfm_cuda_stream(self.h)

    @property
    def launch_count(self): return self.L.sdrjfm_launch_count(self.h)

    def pilot_stats(self):
        """per stream [iterations summed over windows, max per window, fall-back windows, windows]
        of the parallel-in-time pilot PLL solver for the last process call."""
        a = np.zeros((self.n_streams, 4), np.int32)
        self._ck(self.L.sdrjfm_pilot_stats(self.h, a.ctypes.data))
        return a

    def meta(self):
        meta = (Meta * self.n_streams)()
        self._ck(self.L.sdrjfm_get_meta(self.h, C.byref(meta)))
        return [{k: getattr(m, k) for k, _ in Meta._fields_} for m in meta]

    def read_tap(self, name, stream=0, cap=None):
        cap = cap or (self.cfg.max_samples_per_call // self.decim + 16)
        a = np.zeros(cap, _TAP_DTYPE[name])
        n = self.L.sdrjfm_read_tap(self.h, TAP[name], stream, a.ctypes.data, cap)
        if n < 0:
            raise SdrjfmError(n, self.L.sdrjfm_last_error(self.h).decode())
        return a[:n].copy()

    # --- shared tables (rank 0 designs, others import after a broadcast) -----------------
    def tables_export(self):
        n = self.L.sdrjfm_tables_nbytes(self.h)
        buf = (C.c_uint8 * n)()
        self._ck(self.L.sdrjfm_tables_export(self.h, buf, n))
        return np.frombuffer(buf, dtype=np.uint8).copy()

    def tables_import(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._ck(self.L.sdrjfm_tables_import(self.h, blob.ctypes.data, blob.size))
