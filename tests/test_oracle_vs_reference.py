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
