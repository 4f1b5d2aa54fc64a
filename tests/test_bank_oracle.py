"""The bank-builder oracle (oracle/bank_oracle.py) against scipy and against the windowed restatement in
bank_synth.py that made the synthetic banks (CPU only)."""
import numpy as np


def test_bank_oracle_matches_scipy_and_bank_synth(bas):
    from scipy.signal import resample_poly
    from oracle import bank_oracle
    left, _ = bas.bank_synth.synthetic_hrirs(0)
    h = bas.bank_builder.scipy_filter(8)
    assert np.abs(bank_oracle.resample(left[3], 8, h) - resample_poly(left[3], 8, 1)).max() <= 1e-13
    d = bas.bank_synth.delay_differences(left[:5], 8)
    for i in range(5):
        for j in range(i + 1, 5):
            assert abs(d[i, j] - bank_oracle.delay_difference(left[i], left[j], 8, h)) <= 1e-9
    ho = bas.bank_builder.octave_filter(8)
    assert ho.size % 2 == 1 and abs(ho.sum() - 8.0) < 1e-2 and np.allclose(ho, ho[::-1])
    # a pure 3-sample delay reads as +3 with either filter
    for filt in (h, ho):
        assert abs(bank_oracle.delay_difference(left[0], np.roll(left[0], 3), 8, filt) - 3.0) < 1e-3
