"""CPU: pins the plain C++ port (oracle/fm_oracle.cpp) to the reference's own classes
(oracle/_ref, compiled from /root/reference) — BIT-EXACT on every tap and every table.
The reference has no tests or golden vectors of its own, so this, plus the
fixtures in tests/golden/ that were generated from oracle/_ref, is what pins the oracle."""
import numpy as np
import pytest

N1 = 2304000


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


CASES = [
    ("mono", dict(fm_mode=2, volume_db=0.0), "mono_tone", N1 // 2),
    ("stereo_rds", dict(rds_on=1), "stereo_pilot", N1),
    ("input_filter_audio_lp", dict(rds_on=1, input_filter_hz=165000, lf_cutoff_hz=15000),
     "adjacent_interferer", N1 // 3),
    ("pll_decoder", dict(decoder=2), "stereo_pilot", N1 // 4),
    ("cbb_decoder", dict(decoder=4), "stereo_pilot", N1 // 4),
    ("rbb_decoder", dict(decoder=5), "stereo_pilot", N1 // 4),
    ("diff_decoder", dict(decoder=6), "stereo_pilot", N1 // 4),
    ("lo_gain_pano", dict(lo_hz=25000, lgain=0.9, rgain=1.1, fm_mode=1, panorama=150,
                          balance=-30, sound_sel=1, deemph_us=75), "stereo_pilot", N1 // 3),
    ("no_dc_no_automono", dict(dc_remove=0, auto_mono=0, pss_on=0), "stereo_pilot", N1 // 4),
    ("am_decoder", dict(decoder=1, fm_mode=2), "am_tone", N1 // 4),
    ("rate_6M", dict(input_rate=6000000, rds_on=1), "stereo_pilot", N1 // 2),
    ("rate_10M_filter", dict(input_rate=10000000, input_filter_hz=165000), "stereo_pilot", N1 // 2),
    ("rate_2p4M_lo", dict(input_rate=2400000, lo_hz=-30000), "stereo_pilot", N1 // 4),
]


@pytest.mark.parametrize("name,cfg,gen,n", CASES, ids=[c[0] for c in CASES])
def test_port_matches_reference_bit_exact(chainlib, ref_available, signals, name, cfg, gen, n):
    if not ref_available:
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    x = getattr(signals, gen)(n)
    a = chainlib.Chain("ref", **cfg)
    b = chainlib.Chain("orc", **cfg)
    # ragged streaming: the chains must agree across arbitrary call boundaries too
    cuts = [0, 5, 16384 + 5, n // 2 + 7, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        oa, ob = a.process(x[lo:hi]), b.process(x[lo:hi])
        assert oa["n_fm"] == ob["n_fm"] and oa["n_rds24"] == ob["n_rds24"]
        for k in chainlib.Chain.TAPS:
            assert _same(oa[k], ob[k]), f"{name}: tap {k} differs in [{lo},{hi})"
    assert a.meta() == b.meta()


def test_port_tables_bit_exact(chainlib, ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    cfg = dict(input_filter_hz=165000, lf_cutoff_hz=15000)
    a, b = chainlib.Chain("ref", **cfg), chainlib.Chain("orc", **cfg)
    for w in chainlib.DUMP:
        if w == "squelch_iir":          # ref_ only: the port has no squelch (chain_api.h)
            continue
        da, db = a.dump(w), b.dump(w)
        assert da is not None and db is not None, w
        assert _same(da, db), w


def test_decimation_index_contract(chainlib):
    """fm-rate sample m is emitted when input 12m+11 arrives; rds24 q at fm index 8q+7."""
    c = chainlib.Chain("orc", rds_on=1)
    x = np.zeros(12 * 100, np.complex64)
    for n_in, exp_fm in ((11, 0), (1, 1), (12 * 7 + 11, 8), (1, 9)):
        o = c.process(x[:n_in])
        assert o["n_fm"] == exp_fm - getattr(test_decimation_index_contract, "_seen", 0)
        test_decimation_index_contract._seen = exp_fm
    del test_decimation_index_contract._seen
    c = chainlib.Chain("orc", rds_on=1)
    assert c.process(x[:12 * 7])["n_rds24"] == 0
    assert c.process(x[:12])["n_rds24"] == 1
    # impulse: which input index the newest tap touches
    c = chainlib.Chain("orc", dc_remove=0)
    imp = np.zeros(12 * 20, np.complex64)
    imp[12 * 5 + 11] = 1.0
    z = c.process(imp)["fm_z"]
    assert np.all(z[:5] == 0) and z[5] != 0


def test_checker_runtime_setters_are_consistent(chainlib, ref_available, signals):
    """ref_update (the run-time setters of the checker): an update that changes nothing leaves the stream
    bit-identical to uninterrupted processing; settings given at construction and settings applied by an
    update before the first sample are the same chain; setDCRemove zeroes RfDC."""
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    n = N1 // 3
    x = signals.dc_offset(signals.stereo_pilot(n))
    cfg = dict(fm_mode=1, panorama=140, sound_sel=1, balance=-30, deemph_us=75, volume_db=-3.0, decoder=4,
               lgain=0.9, rgain=1.1)
    whole = chainlib.Chain("ref", **cfg).process(x)
    a = chainlib.Chain("ref", **cfg)
    first = a.process(x[:n // 2 + 5])
    a.update(actions=0)
    second = a.process(x[n // 2 + 5:])
    b = chainlib.Chain("ref")
    b.update(actions=0, **cfg)
    late = b.process(x)
    for k in chainlib.Chain.TAPS:
        assert _same(np.concatenate([first[k], second[k]]), whole[k]), k
        assert _same(late[k], whole[k]), k
    c = chainlib.Chain("ref")
    c.process(x[:n // 2])
    assert abs(c.meta()["dc_rf_re"]) > 1e-4
    c.update(actions=4, dc_remove=1)
    assert c.meta()["dc_rf_re"] == 0.0 and c.meta()["dc_rf_im"] == 0.0


def test_checker_lf_spectrum_against_float64_model(chainlib, ref_available):
    """ref_lf_spectrum (ls_scope's arithmetic restated around the reference's Fft_transform) against an
    independent float64 numpy model of ls-scope.cpp:50-52, 76-92, 130-193."""
    if not ref_available:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for N, D, avg, zoom, full in ((2048, 512, 5, 1, False), (2048, 512, 5, 1, True), (1024, 256, 3, 2, True),
                                  (4096, 1024, 1, 4, False)):
        t = np.arange(N * 9 + 100)
        v = (0.5 * np.exp(2j * np.pi * 0.11 * t) + 0.2 * np.cos(2 * np.pi * 0.031 * t)
             + 0.01 * (rng.standard_normal(len(t)) + 1j * rng.standard_normal(len(t)))).astype(np.complex64)
        got = chainlib.ref_lf_spectrum(v, N, D, avg, zoom, full)
        i = np.arange(N)
        win = (0.43 - 0.5 * np.cos(2 * np.pi * i / N) + 0.08 * np.cos(4 * np.pi * i / (N - 1))).astype(np.float32)
        factor = (N // D) // 2
        factor = factor // zoom if factor // zoom >= 1 else 1
        avgbuf = np.zeros(D); disp = np.zeros(D); want = []
        for b in range(len(v) // N):
            f = np.abs(np.fft.fft(v[b * N:(b + 1) * N].astype(np.complex128) * win))
            y = np.zeros(D)
            if full:
                for k in range(D // 2):
                    y[D // 2 + k] = f[k * factor:(k + 1) * factor].mean()
                    y[D // 2 - 1 - k] = f[N - 1 - np.arange(k * factor, (k + 1) * factor)].mean()
            else:
                y = f[:D * factor].reshape(D, factor).mean(axis=1)
            if b == 0:
                avgbuf = y.copy()
            else:
                avgbuf = y / avg + (avg - 1.0) / avg * avgbuf
                disp = avgbuf.copy()
            want.append(disp.copy())
        want = np.array(want)
        assert got.shape == want.shape == (9, D)
        assert np.max(np.abs(got - want)) < 2e-6 * np.max(want)
        assert np.max(want[-1]) > 1.0
