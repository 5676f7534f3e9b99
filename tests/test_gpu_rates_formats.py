"""GPU parity for BASELINE config 4 (device-rate sweep) and for device-native sample formats.

Config 4, reference semantics: at 2.4 / 6 / 10 MS/s the reference's own
constructor arithmetic decimates by 12 / 30 / 48 (stage 1: 25 taps /6, stage 2: IRate/192000+1
taps /(IRate/192000)) and then treats the result as 192 kHz.  The CUDA path runs the same
cascade as one real polyphase FIR (frontend_poly.cuh); the checker is the reference's own
classes constructed with that input rate.  Index contract: fm sample m <-> input D m + D - 1.

Sample formats: the reference's device handlers turn u8 / s8 / s16 into complex float on the CPU
(rtlsdr (b-127)/128, hackrf b/128, sdrplay|pluto|lime|airspy v/2048..8192); the CUDA path reads the
device bytes directly.  Results must equal the float path bit for bit (the conversion is exact).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a.astype(np.complex128)) ** 2)))


@pytest.fixture(scope="module")
def checker(chainlib, ref_available):
    which = "ref" if ref_available else "orc"
    return lambda **cfg: chainlib.Chain(which, **cfg)


TAPS = ("fm_z", "demod", "pilot_phase", "locked", "lr", "audio192")


def run_gpu(pkg, x, input_rate, chunks=None, raw=None, front_end_mode=0, **cfg):
    """x: complex64 [n] (raw=None) or component array [n, 2] with raw=(fmt, denominator)."""
    n = x.shape[0]
    p = pkg.FmProcessorB200(n_streams=1, input_rate=input_rate, max_samples_per_call=max(chunks) if chunks else n,
                            front_end_mode=front_end_mode)
    p.configure(**cfg)
    taps = {k: [] for k in TAPS}
    audio, rds = [], []
    pos = 0
    for c in (chunks or [n]):
        if pos >= n:
            break
        if raw:
            a, r = p.process_raw(x[pos:pos + c], raw[0], raw[1])
        else:
            a, r = p.process(x[pos:pos + c])
        audio.append(a); rds.append(r)
        for k in taps:
            taps[k].append(p.read_tap(k, 0))
        pos += c
    out = {k: np.concatenate(v) for k, v in taps.items()}
    out["audio48"] = np.concatenate(audio, axis=1)[0]
    out["rds24"] = np.concatenate(rds, axis=1)[0]
    p.close()
    return out


@pytest.mark.parametrize("fs,chunks", [
    (2400000, None), (6000000, None), (10000000, None),
    (6000000, [16384, 5, 30 * 7001 + 13, 16384 * 20, 10 ** 7]),
    (10000000, [48 * 3 + 1, 16384, 999999, 10 ** 7]),
])
def test_rate_sweep_matches_reference(pkg, signals, checker, fs, chunks):
    """stereo MPX + pilot at each device rate; the signal's time base is 192000 * D so that the
    chain sees a 19 kHz pilot (the nominal rate enters the tap design, rfDcAlpha and the LO table)."""
    D = pkg.front_end_decimation(fs)
    n = D * 192000 + 7                                     # 1 s of fm-rate output, ragged tail
    x = signals.dc_offset(signals.stereo_pilot(n, fs=192000 * D))
    cfg = dict(fm_mode=0, volume_db=0.0)
    ref = checker(input_rate=fs, **cfg).process(x)
    got = run_gpu(pkg, x, fs, chunks=chunks, **cfg)
    assert len(got["demod"]) == ref["n_fm"] == n // D                   # index contract
    e = rms(got["fm_z"] - ref["fm_z"]) / rms(ref["fm_z"])
    print(fs, "fm_z rel", e, "demod", rms(got["demod"] - ref["demod"]), "audio192", rms(got["audio192"] - ref["audio192"]))
    assert e < 2e-6
    assert rms(got["demod"] - ref["demod"]) < 1e-5
    assert rms(got["audio192"] - ref["audio192"]) < 1e-5
    assert np.array_equal(got["locked"], ref["locked"]) and ref["locked"][-1] == 1
    assert got["audio48"].shape[0] == ref["n_fm"] // 4


@pytest.mark.parametrize("fs,cfg", [
    (6000000, dict(fm_mode=0, input_filter_hz=165000, volume_db=0.0)),
    (10000000, dict(fm_mode=0, input_filter_hz=165000, volume_db=0.0)),
    (2400000, dict(fm_mode=0, input_filter_hz=165000, volume_db=0.0)),
    (6000000, dict(fm_mode=0, lo_hz=30000, lgain=0.9, rgain=1.05, volume_db=0.0)),
    (10000000, dict(fm_mode=0, lo_hz=-47000, input_filter_hz=165000, volume_db=0.0)),
    (6000000, dict(fm_mode=0, rds_on=1, volume_db=-6.0)),
])
def test_rate_sweep_input_filter_lo_rds(pkg, signals, checker, fs, cfg):
    """the optional front-end blocks at the device rates: inputFilter (251 taps designed at the
    device rate, delay 65285 = D q + 5 samples), local oscillator (table of fs entries), RDS."""
    D = pkg.front_end_decimation(fs)
    n = D * 192000
    lo = cfg.get("lo_hz", 0)
    t = np.arange(n) / float(fs)
    x = signals.batch_stream(2, n, fs=192000 * D) * np.exp(2j * np.pi * lo * t)
    x = signals.dc_offset(x.astype(np.complex64))
    ref = checker(input_rate=fs, **cfg).process(x)
    got = run_gpu(pkg, x, fs, chunks=[n // 3 + 7, 16384, n], **cfg)
    assert len(got["demod"]) == ref["n_fm"]
    e = rms(got["fm_z"] - ref["fm_z"]) / rms(ref["fm_z"])
    print(fs, cfg, "fm_z rel", e, "demod", rms(got["demod"] - ref["demod"]), "audio192", rms(got["audio192"] - ref["audio192"]))
    assert e < (2e-5 if lo else 6e-6)      # includes the float32 rounding of the reference's 65536-point FFT filter
    assert rms(got["demod"] - ref["demod"]) < 2e-5
    assert rms(got["audio192"] - ref["audio192"]) < 1e-5
    if cfg.get("rds_on"):
        assert got["rds24"].shape[0] == ref["n_rds24"]
        assert rms(got["rds24"] - ref["rds24"]) < 1e-5


def _quantise(x, fmt):
    """what the device would have delivered for the analogue signal x, and the floats the
    reference's handler makes of those bytes."""
    z = np.stack([x.real, x.imag], axis=-1).astype(np.float64)
    if fmt == "u8":
        b = np.clip(np.round(z * 128.0 + 127.0), 0, 255).astype(np.uint8)
        f = (b.astype(np.float32) - 127.0) / 128.0           # rtlsdr-handler.cpp:291-292
        den = 128
    elif fmt == "s8":
        b = np.clip(np.round(z * 128.0), -128, 127).astype(np.int8)
        f = b.astype(np.float32) / 128.0                     # hackrf-handler.cpp:364-365
        den = 128
    else:
        den = int(fmt.split("/")[1])
        b = np.clip(np.round(z * den), -32768, 32767).astype(np.int16)
        f = b.astype(np.float32) / np.float32(den)           # sdrplay-handler.cpp:486-487
    return b, (f[..., 0] + 1j * f[..., 1]).astype(np.complex64), den


@pytest.mark.parametrize("fmt,fs", [("u8", 2304000), ("s8", 2304000), ("s16/2048", 2304000),
                                    ("s16/4096", 6000000), ("s16/8192", 10000000), ("u8", 2400000)])
def test_device_sample_formats_equal_handler_conversion(pkg, signals, checker, fmt, fs):
    D = pkg.front_end_decimation(fs)
    n = D * 96000 + 5
    x = signals.dc_offset(signals.stereo_pilot(n, fs=192000 * D, amp=0.7))
    raw, xf, den = _quantise(x, fmt)
    cfg = dict(fm_mode=0, volume_db=0.0)
    chunks = [16384, 7, n // 2, n]
    a = run_gpu(pkg, raw, fs, chunks=chunks, raw=(fmt.split("/")[0], den), **cfg)
    b = run_gpu(pkg, xf, fs, chunks=chunks, **cfg)
    for k in ("fm_z", "demod", "audio192"):
        if D in (12, 48):
            # float path = K1t (TMA), raw path = K1g: same taps, different summation order; behind the
            # discriminator a last-bit difference can flip an atan-table index (7.9e-5 per flip)
            assert rms(a[k] - b[k]) < (1e-6 if k == "fm_z" else 5e-6), k
        else:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    ref = checker(input_rate=fs, **cfg).process(xf)
    assert rms(a["audio192"] - ref["audio192"]) < 1e-5
    assert rms(a["demod"] - ref["demod"]) < 1e-5


def test_generic_front_end_equals_tuned_kernel(pkg, signals, checker, monkeypatch):
    """SDRJFM_GENERIC_FE=1 routes the reference's own rate / float format through K1g as well."""
    n = 2304000 // 2
    x = signals.dc_offset(signals.stereo_pilot(n))
    for cfg in (dict(fm_mode=0, volume_db=0.0), dict(fm_mode=0, input_filter_hz=165000, volume_db=0.0),
                dict(fm_mode=0, lo_hz=30000, volume_db=0.0)):
        a = run_gpu(pkg, x, 2304000, chunks=[16384 * 3 + 5, n], **cfg)
        monkeypatch.setenv("SDRJFM_GENERIC_FE", "1")
        b = run_gpu(pkg, x, 2304000, chunks=[16384 * 3 + 5, n], **cfg)
        monkeypatch.delenv("SDRJFM_GENERIC_FE")
        ref = checker(**cfg).process(x)
        assert rms(a["fm_z"] - b["fm_z"]) / rms(a["fm_z"]) < 1e-6
        assert rms(b["audio192"] - ref["audio192"]) < 1e-5


# ---- front_end_mode 1: rational polyphase resampler (new block; float64 model of the same taps) ----
def resampler_model(x, fs, tables, dc_remove=True):
    """float64 model of csrc/resample.cuh on the device's own float32 taps: per-sample RF DC removal
    as in fm-processor.cpp:423-446, stage A (49 taps, /5, A[k] <-> input 5k+4), stage B (L/M polyphase,
    y[m] = sum_j hB[phi_m][j] A[n_m - j], n_m = floor(m M / L), phi_m = (m M) mod L)."""
    from scipy.signal import lfilter
    L, M, P, hA, hB = tables.resampler
    y = x.astype(np.complex128)
    if dc_remove:
        alpha = float(np.float32(1.0) / np.float32(fs))
        r = lfilter([alpha], [1.0, -(1.0 - alpha)], y)
        y = y - np.clip(r.real, -0.01, 0.01) - 1j * np.clip(r.imag, -0.01, 0.01)
    A = np.convolve(y, hA.astype(np.float64))[4::5][:len(y) // 5]
    n_out = -((-len(A) * L) // M)
    t = np.arange(n_out, dtype=np.int64) * M
    n, phi = t // L, t % L
    Ap = np.concatenate([np.zeros(P - 1, np.complex128), A])
    z = np.zeros(n_out, np.complex128)
    hB = hB.astype(np.float64)
    for j in range(P):
        z += hB[phi, j] * Ap[n - j + P - 1]
    return z


@pytest.mark.parametrize("fs,chunks", [
    (2400000, None), (6000000, None), (10000000, None), (2304000, None),
    (6000000, [16384, 3, 1, 7001 * 5 + 2, 16384 * 20, 10 ** 7]),
    (10000000, [16384] * 40 + [10 ** 7]),
])
def test_polyphase_resampler_mode_matches_float64_model(pkg, signals, chainlib, ref_available, fs, chunks):
    """config 4 proper: 2.4 / 6 / 10 MS/s resampled to exactly 192 kHz (L/M = 2/25, 4/125, 12/625 overall).
    Front end against the float64 model of the same taps; the chain behind it against the reference's
    classes entered at the fm rate (the reference's own decimator bypass, fm-processor.cpp:471)."""
    n = fs // 2 + 3
    x = signals.dc_offset(signals.stereo_pilot(n, fs=fs))
    cfg = dict(fm_mode=0, volume_db=0.0)
    T = pkg.design_tables(input_rate=fs)
    L, M, P, hA, hB = T.resampler
    assert L * fs == 5 * 192000 * M and abs(hA.sum() - 1) < 1e-6 and np.allclose(hB.sum(axis=1), 1, atol=1e-6)
    z = resampler_model(x, fs, T)
    p = pkg.FmProcessorB200(n_streams=1, input_rate=fs, max_samples_per_call=max(chunks) if chunks else n,
                            front_end_mode=1)
    p.configure(**cfg)
    taps = {k: [] for k in TAPS}
    pos, n48 = 0, 0
    for c in (chunks or [n]):
        if pos >= n:
            break
        a, _ = p.process(x[pos:pos + c])
        n48 += a.shape[1]
        for k in taps:
            taps[k].append(p.read_tap(k, 0))
        pos += c
    p.close()
    got = {k: np.concatenate(v) for k, v in taps.items()}
    assert len(got["fm_z"]) == len(z) == -((-(n // 5) * L) // M)       # index contract of the resampler
    assert n48 == len(z) // 4
    e = rms(got["fm_z"] - z) / rms(z)
    which = "ref" if ref_available else "orc"
    ref = chainlib.Chain(which, input_rate=192000, dc_remove=0, **cfg).process_fm(z.astype(np.complex64))
    print(fs, "L/M", L, M, "P", P, "fm_z rel", e, "demod", rms(got["demod"] - ref["demod"]),
          "audio192", rms(got["audio192"] - ref["audio192"]))
    assert e < 1e-5
    assert rms(got["demod"] - ref["demod"]) < 1e-5
    assert rms(got["audio192"] - ref["audio192"]) < 1e-5
    assert np.array_equal(got["locked"], ref["locked"])
    # the resampled stream really is at 192 kHz: the pilot sits at 19 kHz
    d = got["demod"][-65536:].astype(np.float64)
    spec = np.abs(np.fft.rfft(d * np.hanning(len(d))))
    k0, k1 = int(17000 * len(d) / 192000), int(21000 * len(d) / 192000)
    f = (k0 + np.argmax(spec[k0:k1])) * 192000.0 / len(d)
    assert abs(f - 19000.0) < 3.0


# ---- airspy: native-rate int16 with the handler's per-millisecond linear interpolation ----------
def airspy_handler_model(raw, native_rate):
    """restatement of airspyHandler::data_available + the map tables of its constructor
    (devices/airspy/airspy-handler.cpp:117-128, 283-309) in float32: raw int16 [n, 2] at the native
    rate -> complex64 at 2 304 000 samples/s.  (The handler itself needs libairspy and Qt: this
    model is the checker for the conversion; everything behind it is checked against the reference.)"""
    B = native_rate // 1000
    x = raw.astype(np.float32) / np.float32(2048)
    j = np.arange(2304)
    pos = j * (float(np.float32(B)) / 2304.0)
    mi = np.floor(pos).astype(np.int64)
    mf = (pos - mi).astype(np.float32)
    blocks = (len(raw) - 1) // B
    out = np.empty((blocks * 2304, 2), np.float32)
    one_minus = (np.float32(1) - mf)
    for b in range(blocks):
        x0, x1 = x[b * B + mi], x[b * B + mi + 1]
        out[b * 2304:(b + 1) * 2304] = x1 * mf[:, None] + x0 * one_minus[:, None]
    return (out[:, 0] + 1j * out[:, 1]).astype(np.complex64)


@pytest.mark.parametrize("native,chunks", [(2500000, None), (3000000, [16384, 5, 2999, 3001, 250001, 10 ** 7]),
                                           (6000000, [65536] * 30 + [10 ** 7])])
def test_airspy_native_rate_conversion(pkg, signals, checker, monkeypatch, native, chunks):
    n = native // 2 + 17                                   # 0.5 s of native samples, ragged
    x = signals.dc_offset(signals.stereo_pilot(n, fs=native, amp=0.6))
    raw = np.clip(np.round(np.stack([x.real, x.imag], -1).astype(np.float64) * 2048), -32768, 32767).astype(np.int16)
    xf = airspy_handler_model(raw, native)
    assert len(xf) == ((n - 1) // (native // 1000)) * 2304
    cfg = dict(fm_mode=0, volume_db=0.0)
    p = pkg.FmProcessorB200(n_streams=1, input_rate=2304000, max_samples_per_call=max(chunks) if chunks else n + 8192)
    p.configure(**cfg)
    p.set_nativeRate(native)
    taps = {k: [] for k in TAPS}
    pos = 0
    for c in (chunks or [n]):
        if pos >= n:
            break
        p.process_raw(raw[pos:pos + c], "airspy", 2048)
        for k in taps:
            taps[k].append(p.read_tap(k, 0))
        pos += c
    p.close()
    a = {k: np.concatenate(v) for k, v in taps.items()}
    # the same floats through the same (generic) front-end kernel: bit-identical
    monkeypatch.setenv("SDRJFM_GENERIC_FE", "1")
    b = run_gpu(pkg, xf, 2304000, **cfg)
    monkeypatch.delenv("SDRJFM_GENERIC_FE")
    assert len(a["demod"]) == len(b["demod"]) == len(xf) // 12
    assert np.array_equal(a["fm_z"].view(np.uint32), b["fm_z"].view(np.uint32))
    # (behind the front end the one-pole scans depend on how the stream was cut into calls: rounding only)
    assert rms(a["demod"] - b["demod"]) < 1e-6 and rms(a["audio192"] - b["audio192"]) < 1e-6
    ref = checker(**cfg).process(xf)
    print(native, "audio192 vs reference on the handler's floats", rms(a["audio192"] - ref["audio192"]))
    assert rms(a["audio192"] - ref["audio192"]) < 1e-5 and rms(a["demod"] - ref["demod"]) < 1e-5
    assert np.array_equal(a["locked"], ref["locked"])


@pytest.mark.parametrize("name,rate,mode,cfg,raw", [
    ("input_filter", 2304000, 0, dict(fm_mode=0, rds_on=1, input_filter_hz=165000), None),
    ("local_oscillator", 2304000, 0, dict(fm_mode=0, rds_on=1, lo_hz=30000), None),
    ("rate_6M", 6000000, 0, dict(fm_mode=0, rds_on=1), None),
    ("resampler_2p4M", 2400000, 1, dict(fm_mode=0, rds_on=1), None),
    ("u8", 2304000, 0, dict(fm_mode=0, rds_on=1), ("u8", 128)),
    ("squelch_audio_lp", 2304000, 0, dict(fm_mode=0, rds_on=1, squelch_mode=1, squelch_value=50, lf_cutoff_hz=15000), None),
    ("rds_mode_2", 2304000, 0, dict(fm_mode=0, rds_on=2), None),
    ("rds_mode_3", 2304000, 0, dict(fm_mode=0, rds_on=3), None),
])
def test_copies_of_a_stream_are_bit_identical_across_lanes(pkg, signals, name, rate, mode, cfg, raw):
    """determinism under concurrency: 130 streams (4 lanes, RDS side streams) carrying two distinct
    signals in a shuffled order, two calls; every copy of a signal must give bit-identical audio and
    RDS output wherever it runs (a race between a lane's kernels and the others' shows up here, not
    in a tolerance test)."""
    S = 130
    D = (rate // 192000) if mode == 1 else pkg.front_end_decimation(rate)
    n = D * 48000 + 5 * D
    fs = rate if mode == 1 else 192000 * D
    base = [signals.batch_stream(60 + k, 2 * n, fs=fs) for k in range(2)]
    order = np.random.default_rng(11).permutation(S) % 2
    p = pkg.FmProcessorB200(n_streams=S, input_rate=rate, max_samples_per_call=n, keep_taps=False, front_end_mode=mode)
    p.configure(volume_db=-6.0, **cfg)
    p.setRdsSymbolStage(True)
    outs = []
    for c in range(2):
        x = np.stack([base[k][c * n:(c + 1) * n] for k in order])
        if raw:
            q, _, den = _quantise(x, raw[0])
            a, r = p.process_raw(q, raw[0], den)
        else:
            a, r = p.process(x)
        bits = [p.read_rds_bits(s) for s in range(S)]
        outs.append((a, r, bits))
    p.close()
    for a, r, bits in outs:
        for k in range(2):
            idx = np.nonzero(order == k)[0]
            for i in idx[1:]:
                assert np.array_equal(a[i].view(np.uint64), a[idx[0]].view(np.uint64)), (name, k, i)
                assert np.array_equal(r[i].view(np.uint64), r[idx[0]].view(np.uint64)), (name, k, i)
                assert np.array_equal(bits[i], bits[idx[0]]), (name, k, i)
        assert not np.array_equal(a[np.nonzero(order == 0)[0][0]], a[np.nonzero(order == 1)[0][0]])


@pytest.mark.parametrize("fs,chunks", [
    (2048000, None), (1920000, [16384, 5, 6 * 70001 + 3, 10 ** 7]), (3200000, None),
    (4000000, [18 * 7 + 1, 16384 * 9, 10 ** 7]), (8000000, None), (12000000, [16384, 10 ** 8]),
])
def test_any_input_rate_of_the_reference_arithmetic(pkg, signals, checker, fs, chunks):
    """The reference takes whatever rate the WAV file carries (devices/filereader/filehulp.cpp:62) through its
    constructor arithmetic (fm-processor.cpp:36,68-75): stage 1 = 25 taps / 6, stage 2 = D2 + 1 taps / D2 with
    D2 = (inputRate / 6) / 192000.  Rates without a tuned kernel (any D2 from 1 to 10 other than 2, 5, 8: e.g. the
    colibri's 1.92 MS/s, rtlsdr's 2.048 MS/s, 4, 8, 12 MS/s) run the reference-order front end with D2 at run time:
    fm-rate samples bit-identical to the reference's, exact output counts, audio within the north-star tolerance."""
    D = pkg.front_end_decimation(fs)
    n = D * 48000 + 7
    x = signals.dc_offset(signals.stereo_pilot(n, fs=192000 * D, amp=0.6))
    cfg = dict(fm_mode=0, volume_db=0.0)
    ref = checker(input_rate=fs, **cfg).process(x)
    got = run_gpu(pkg, x, fs, chunks=chunks, **cfg)
    assert len(got["fm_z"]) == ref["n_fm"] == n // D
    tuned = D in (12, 30, 48)
    e = rms(got["fm_z"] - ref["fm_z"]) / rms(ref["fm_z"])
    print(fs, "D", D, "fm_z rel", e, "demod", rms(got["demod"] - ref["demod"]), "audio192", rms(got["audio192"] - ref["audio192"]))
    if tuned:
        assert e < 2e-6
    else:
        assert np.array_equal(got["fm_z"].view(np.uint32), ref["fm_z"].view(np.uint32))
    assert rms(got["demod"] - ref["demod"]) < 1e-5 and rms(got["audio192"] - ref["audio192"]) < 1e-5
    assert np.array_equal(got["locked"], ref["locked"])


@pytest.mark.parametrize("fmt,fs", [("u8", 2304000), ("s16/4096", 6000000), ("s8", 3200000)])
def test_reference_order_front_end_reads_device_formats(pkg, signals, checker, fmt, fs):
    """front_end_mode 2 (K1x: the reference's per-sample DC recurrence and tap-by-tap decimators) on the device's own
    bytes: bit-identical to the same stream handed over as the handler's floats, and fm_z bit-identical to the
    reference's classes fed with those floats (3.2 MS/s is a rate only this front end serves)."""
    D = pkg.front_end_decimation(fs)
    n = D * 60000 + 5
    x = signals.dc_offset(signals.stereo_pilot(n, fs=192000 * D, amp=0.7))
    raw, xf, den = _quantise(x, fmt)
    cfg = dict(fm_mode=0, volume_db=0.0)
    chunks = [16384, 7, n // 2, n]
    a = run_gpu(pkg, raw, fs, chunks=chunks, raw=(fmt.split("/")[0], den), front_end_mode=2, **cfg)
    b = run_gpu(pkg, xf, fs, chunks=chunks, front_end_mode=2, **cfg)
    for k in ("fm_z", "demod", "audio192"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    ref = checker(input_rate=fs, **cfg).process(xf)
    m = min(len(a["fm_z"]), len(ref["fm_z"]))
    assert m == len(ref["fm_z"]) and np.array_equal(a["fm_z"][:m].view(np.uint32), ref["fm_z"][:m].view(np.uint32))
    assert rms(a["audio192"] - ref["audio192"]) < 1e-5
