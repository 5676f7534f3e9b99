"""CPU: the C-ABI library loads and exports every symbol include/sdrjfm_b200.h declares,
fails loudly without a device, and its host-designed tables are bit-identical to what the
reference's constructors build (oracle/_ref dumps)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(pkg.HEADER).read()
    names = sorted(set(re.findall(r"\b(sdrjfm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = pkg.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.sdrjfm_version()


def test_host_mirror_binds_every_entry_point_and_reference_setter(pkg):
    """the Python mirror of the reference interface reaches every C entry point, and carries the
    fmProcessor setter names (includes/fm/fm-processor.h:122-157) the parity tests are written in."""
    hdr = open(pkg.HEADER).read()
    src = open(pkg.__file__).read()
    names = sorted(set(re.findall(r"\b(sdrjfm_[a-z0-9_]+)\s*\(", hdr)))
    unbound = [n for n in names if n not in src]
    assert not unbound, unbound
    for setter in ("setfmMode", "setFMdecoder", "setSoundMode", "setStereoPanorama", "setSoundBalance",
                   "setDeemphasis", "setVolume", "setlfcutoff", "setBandwidth", "setAttenuation",
                   "setfmRdsSelector", "set_localOscillator", "set_squelchMode", "set_squelchValue",
                   "setAutoMonoMode", "setPSSMode", "setDCRemove", "triggerFrequencyChange",
                   "restartPssAnalyzer", "startScanning", "stopScanning", "setlfPlotType",
                   "setlfPlotZoomFactor"):
        assert hasattr(pkg.FmProcessorB200, setter), setter
    # enum values follow the reference's declarations
    assert pkg.LF_PLOT["RDS_DEMOD"] == 9 and pkg.LF_PLOT["OFF"] == 0
    assert pkg.IQ_FORMAT["cf32"] == 0


def test_create_fails_loudly_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(pkg.SdrjfmError) as e:
        pkg.FmProcessorB200(n_streams=1)
    assert e.value.status == pkg.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_bad_arguments(pkg):
    L = pkg.lib()
    st = C.c_int(0)
    assert not L.sdrjfm_create(None, C.byref(st)) and st.value == pkg.ERR_ARG
    cfg = pkg.Config(1000000, 192000, 48000, 48000, 1, 0, 1000, 0, 0)   # below 6 x fmRate: the reference's stage 1 alone would undershoot the fm rate
    assert not L.sdrjfm_create(C.byref(cfg), C.byref(st)) and st.value == pkg.ERR_UNSUPPORTED
    assert L.sdrjfm_design_tables(100, 192000, 0, 0, None, 0) == pkg.ERR_ARG


def test_tables_match_reference_constructors(pkg, chainlib, ref_available):
    which = "ref" if ref_available else "orc"
    T = pkg.design_tables(input_filter_hz=165000, audio_lp_hz=15000)
    c = chainlib.Chain(which, input_filter_hz=165000, lf_cutoff_hz=15000)
    assert _same(T.fmband1, c.dump("fmband1"))
    assert _same(T.fmband2, c.dump("fmband2"))
    assert _same(T.rdsdecim, c.dump("rdsdecim"))
    assert _same(T.sincos, c.dump("sincos"))
    assert _same(T.atan, c.dump("atan").view(np.float32))
    assert _same(T.pss_lp_freq, c.dump("pss_lp_freq"))
    assert _same(T.rds_bp_freq, c.dump("rds_bp_freq"))
    assert _same(T.audio_lp_freq, c.dump("audio_lp_freq"))
    k = c.dump("consts").view(np.float32)
    assert T.consts[4] == k[0]                       # K_FM
    # the 251 input taps are the real part of the low-pass whose spectrum the reference holds
    spec = np.fft.fft(np.concatenate([T.input_taps.astype(np.float64), np.zeros(65536 - 251)]))
    ref = c.dump("input_filter_freq").astype(np.complex128)
    assert np.abs(spec - ref).max() < 2e-5


def test_composite_equals_cascade(pkg):
    """the 37 real taps x constant gain reproduce the two reference kernels in cascade."""
    T = pkg.design_tables()
    k1, k2 = T.fmband1.astype(np.complex128), T.fmband2.astype(np.complex128)
    casc = np.zeros(37, complex)
    for j in range(3):
        casc[6 * j:6 * j + 25] += k2[j] * k1
    G = complex(T.consts[2], T.consts[3])
    # the stored taps are C' = C + alpha*g (RF DC removal folded in, g = tail sums of C)
    alpha = float(np.float32(1.0) / np.float32(2304000))
    C = (casc / G).real
    g = np.array([C[i + 1:].sum() for i in range(37)])
    comp = (T.composite.astype(np.float64) - alpha * g) * G
    assert np.abs(comp - casc).max() < 3e-7 * np.abs(casc).max()
    assert abs(T.consts[0] - C.sum()) < 1e-6
    assert abs(T.consts[1] - T.composite.astype(np.float64).sum()) < 1e-6


@pytest.mark.parametrize("fs", [2400000, 6000000, 10000000])
def test_tables_at_device_rates_match_reference_constructors(pkg, chainlib, ref_available, fs):
    """BASELINE config 4: at 2.4 / 6 / 10 MS/s the reference's constructor arithmetic gives stage 2
    = IRate/192000 + 1 taps, /(IRate/192000); the host designer must build the same kernels."""
    which = "ref" if ref_available else "orc"
    T = pkg.design_tables(input_rate=fs)
    c = chainlib.Chain(which, input_rate=fs)
    assert _same(T.fmband1, c.dump("fmband1"))
    assert _same(T.fmband2, c.dump("fmband2"))
    d = pkg.front_end_decimation(fs)
    assert d == int(T.hdr["decim1"]) * int(T.hdr["decim2"]) == {2400000: 12, 6000000: 30, 10000000: 48}[fs]
    assert int(T.hdr["ncomp"]) == 25 + 6 * (d // 6)


def test_squelch_filter_design_matches_reference(pkg, chainlib, ref_available):
    """the two 20th-order Chebyshev IIRs of the squelch (iir-filters.cpp via squelchClass.cpp:12-21):
    host-side restatement of the design, bit-identical coefficients and gains."""
    if not ref_available:
        pytest.skip("oracle/_ref not available")
    ref = chainlib.Chain("ref").dump("squelch_iir").view(np.float32)[:82]
    mine = pkg.design_tables().squelch_iir
    assert _same(np.ascontiguousarray(ref), np.ascontiguousarray(mine))


def test_rds_symbol_stage_tables_match_reference(pkg, chainlib, ref_available):
    """matched filter, low-pass and the 8-biquad band-pass of rdsDecoder_1 (rds-decoder-1.cpp:45-112,
    iir-filters.cpp:315-382,555-596): host-side restatement, bit-identical to the reference's objects."""
    if not ref_available:
        pytest.skip("oracle/_ref not available")
    r = chainlib.Rds1()
    match, lp, bp = pkg.design_tables().rds_symbol
    for mine, which in ((match, 0), (lp, 1), (bp, 2)):
        assert _same(r.dump(which), np.ascontiguousarray(mine))


def test_rds2_matched_filter_is_bit_identical(pkg, chainlib, ref_available):
    """the 45-tap root-raised-cosine matched filter of rdsDecoder_2 (rds-decoder-2.cpp:67-71, shaping_filter.cpp:4-56)
    restated on the host must be the reference's bit for bit."""
    if not ref_available:
        pytest.skip("oracle/_ref not available")
    mine = pkg.design_aux(0, 24000)
    ref = chainlib.Rds2(24000).matched_filter()
    assert len(mine) == len(ref) == 45
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_test_tone_table_follows_the_reference_state_machine(pkg, chainlib, ref_available):
    """insertTestTone (fm-processor.cpp:800-823) restated in the checker: fed with zeros, the burst it inserts must be
    the table the library indexes by cycle position, and it must start where the table says."""
    if not ref_available:
        pytest.skip("oracle/_ref not available")
    t = pkg.design_aux(1, 48000)
    arm, burst = int(t[0]), t[1:]
    assert arm == 96001 and len(burst) == 1200
    r = chainlib.RefPost(48000)
    r.set(1, 0)
    out, _ = r.process(np.zeros(2 * (arm + len(burst)) + 10, np.complex64))
    for cyc in (0, 1):
        a = cyc * (arm + len(burst))
        assert not out[a:a + arm].any()
        seg = out[a + arm:a + arm + len(burst)]
        assert np.array_equal(seg.real.view(np.uint32), (np.float32(0.9) * burst).view(np.uint32))
        assert np.array_equal(seg.real, seg.imag)


def test_second_converter_design(pkg):
    """rational L / M, unit DC gain per polyphase branch, for the rates main.cpp / the ini file can ask for"""
    for rate, L, M in ((192000, 4, 1), (44100, 147, 160), (96000, 2, 1), (32000, 2, 3)):
        t = pkg.design_aux(2, 48000, rate)
        assert (int(t[0]), int(t[1])) == (L, M)
        h = t[2:].reshape(32, L).astype(np.float64)          # h [j * L + p]
        assert np.allclose(h.sum(axis=0), 1.0, atol=1e-6)


def test_rds_group_bits_pass_the_reference_synchroniser(signals, chainlib, ref_available):
    """the test signal of the RDS_3 symbol stage carries checkwords the reference's own rdsBlockSynchronizer accepts
    (IEC 62106 generator polynomial and offset words, version A and B groups): every group comes back, in order."""
    if not ref_available:
        pytest.skip("oracle/_ref not available")
    rng = np.random.default_rng(8)
    groups = [(0xABCD, ((i % 16) << 12) | ((i & 1) << 11) | (i & 0x1F), int(rng.integers(0, 65536)), int(rng.integers(0, 65536)))
              for i in range(25)]
    bits = np.concatenate([rng.integers(0, 2, 41).astype(np.uint8), signals.rds_group_bits(groups)])
    got = chainlib.ref_blocksync_groups(bits)
    assert [tuple(int(v) for v in g) for g in got] == groups
    # one flipped payload bit: the synchroniser drops that group (the burst correction's verdict is not used) and re-locks
    bad = bits.copy(); bad[41 + 104 * 10 + 30] ^= 1
    got = chainlib.ref_blocksync_groups(bad)
    assert 20 <= len(got) < 25 and all(tuple(int(v) for v in g) in groups for g in got)
