"""Seeded synthetic IQ for the five BASELINE.json configs.

The reference has no test signals of its own; these are the inputs both
the CUDA path and the CPU checkers are fed.  All phase arithmetic is float64, the IQ that
leaves here is complex64 — the format `deviceHandler::getSamples` delivers
(devices/device-handler.h:72-75).

MPX convention the reference decodes as (L, R): pilot sin(theta), sub-carrier
+(L-R)*sin(2*theta).
"""
import numpy as np

INPUT_RATE = 2304000
FM_RATE = 192000
PILOT_HZ = 19000.0
DEVIATION_HZ = 75000.0


def _awgn(rng, n, amp, snr_db):
    if snr_db is None:
        return 0.0
    sigma = amp / np.sqrt(2.0) * 10.0 ** (-snr_db / 20.0)
    return sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def fm_modulate(mpx, fs=INPUT_RATE, deviation=DEVIATION_HZ, amp=0.5, f_offset=0.0,
                phase0=0.0):
    """x[n] = amp * exp(j*phi[n]), phi[n] = phi[n-1] + 2*pi*(deviation*mpx[n] + f_offset)/fs."""
    phi = phase0 + np.cumsum(2.0 * np.pi * (deviation * mpx + f_offset) / fs)
    return amp * np.exp(1j * phi)


def mono_tone(n, fs=INPUT_RATE, tone_hz=1000.0, amp=0.5, snr_db=40.0, seed=1234):
    """config 1: single 1 kHz tone, 75 kHz deviation, optional AWGN."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    x = fm_modulate(np.sin(2 * np.pi * tone_hz * t), fs, amp=amp)
    return (x + _awgn(rng, n, amp, snr_db)).astype(np.complex64)


def stereo_mpx(n, fs=INPUT_RATE, left_hz=1000.0, right_hz=None, pilot=0.10, rds_bits=None,
               rds_level=0.05, t0=0.0):
    """0.45(L+R) + 0.45(L-R) sin 2θ + pilot·sin θ (+ 57 kHz BPSK RDS, differential Manchester)."""
    t = t0 + np.arange(n, dtype=np.float64) / fs
    th = 2 * np.pi * PILOT_HZ * t
    L = np.sin(2 * np.pi * left_hz * t) if left_hz else np.zeros(n)
    R = np.sin(2 * np.pi * right_hz * t) if right_hz else np.zeros(n)
    mpx = 0.45 * (L + R) + 0.45 * (L - R) * np.sin(2 * th) + pilot * np.sin(th)
    if rds_bits is not None:
        # 1187.5 bit/s = 57000/48; biphase symbol, differentially encoded
        bit_idx = np.floor(t * 1187.5).astype(np.int64)
        half = (np.floor(t * 2375.0).astype(np.int64) & 1)
        d = np.cumsum(np.asarray(rds_bits, dtype=np.int64)) & 1
        sym = 1.0 - 2.0 * d[bit_idx % len(d)]
        mpx = mpx + rds_level * sym * (1.0 - 2.0 * half) * np.sin(3 * th)
    return mpx


def stereo_pilot(n, fs=INPUT_RATE, left_hz=1000.0, right_hz=None, amp=0.5, snr_db=40.0,
                 seed=1235, pilot=0.10):
    """config 2: stereo MPX with 10 % pilot; L-only tone by default (separation check)."""
    rng = np.random.default_rng(seed)
    x = fm_modulate(stereo_mpx(n, fs, left_hz, right_hz, pilot), fs, amp=amp)
    return (x + _awgn(rng, n, amp, snr_db)).astype(np.complex64)


def adjacent_interferer(n, fs=INPUT_RATE, seed=1236, amp=0.05, snr_db=40.0):
    """config 3: config-2 signal + carrier at +200 kHz, +20 dB, 400 Hz tone."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    want = fm_modulate(stereo_mpx(n, fs), fs, amp=amp)
    adj = fm_modulate(np.sin(2 * np.pi * 400.0 * t), fs, amp=10.0 * amp, f_offset=200000.0)
    return (want + adj + _awgn(rng, n, amp, snr_db)).astype(np.complex64)


RDS_OFFSET_WORDS = dict(A=0x0FC, B=0x198, C=0x168, C2=0x350, D=0x1B4)      # IEC 62106 annex A


def rds_checkword(data16, offset):
    """10-bit checkword of one RDS block: remainder of data x^10 modulo g(x) = x^10 + x^8 + x^7 + x^5 + x^4 + x^3 + 1,
    plus the block's offset word."""
    reg = (int(data16) & 0xFFFF) << 10
    for k in range(25, 9, -1):
        if reg & (1 << k):
            reg ^= 0x5B9 << (k - 10)
    return (reg & 0x3FF) ^ offset


def rds_group_bits(groups):
    """groups: iterable of (A, B, C, D) 16-bit block contents -> the 104 bits per group, msb first, with checkwords
    (block C takes offset C' in version-B groups, i.e. when bit 11 of block B is set)."""
    out = []
    for a, b, c, d in groups:
        offs = (RDS_OFFSET_WORDS["A"], RDS_OFFSET_WORDS["B"],
                RDS_OFFSET_WORDS["C2"] if (b >> 11) & 1 else RDS_OFFSET_WORDS["C"], RDS_OFFSET_WORDS["D"])
        for w, o in zip((a, b, c, d), offs):
            word = ((int(w) & 0xFFFF) << 10) | rds_checkword(w, o)
            out.extend((word >> k) & 1 for k in range(25, -1, -1))
    return np.array(out, dtype=np.uint8)


def batch_stream(s, n, fs=INPUT_RATE, amp=0.5, snr_db=40.0, with_rds=True, rds_bits=None):
    """config 5, stream s: config-2 MPX with tone 400 + 10·(s mod 256) Hz (the 256 streams of the config; beyond them the
    tones repeat instead of climbing into the pilot), RDS bits from rng(2000+s) (or `rds_bits`, repeated)."""
    rng = np.random.default_rng(2000 + s)
    bits = rng.integers(0, 2, size=4096) if with_rds else None
    if rds_bits is not None:
        bits = np.asarray(rds_bits)
    mpx = stereo_mpx(n, fs, left_hz=400.0 + 10.0 * (s % 256), right_hz=None, rds_bits=bits)
    x = fm_modulate(mpx, fs, amp=amp, phase0=0.1 * s)
    return (x + _awgn(rng, n, amp, snr_db)).astype(np.complex64)


def dc_offset(x, dc=0.004 + 0.003j):
    """adds a front-end DC offset (exercises the RF DC remover, fm-processor.cpp:423-446)."""
    return (x + np.complex64(dc)).astype(np.complex64)


def am_tone(n, fs=INPUT_RATE, tone_hz=1000.0, depth=0.5, amp=0.4, f_offset=3000.0, snr_db=40.0, seed=1237):
    """AM carrier (for the AM decoder, fm-demodulator.cpp:215-241): amp (1 + depth sin) at a small offset."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    x = amp * (1.0 + depth * np.sin(2 * np.pi * tone_hz * t)) * np.exp(2j * np.pi * f_offset * t)
    return (x + _awgn(rng, n, amp, snr_db)).astype(np.complex64)
