"""GPU bank builder (upsample_irs.m on the device) against its float64 numpy oracle
(oracle/bank_oracle.py) and against the scipy-based restatement that made the synthetic banks of the
other tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind', ['octave', 'scipy'])
def test_small_bank_vs_numpy_twin(bas, kind):
    from oracle import bank_oracle
    bb = bas.bank_builder
    rng = np.random.default_rng(17)
    n_rows, n, u = 9, 48, 8
    t = np.arange(n)[None, :] - (12 + 6 * rng.uniform(-1, 1, n_rows))[:, None]
    left = np.sinc(0.8 * t) * np.exp(-0.02 * t ** 2) + 0.05 * rng.standard_normal((n_rows, n))
    right = left[::-1] * 0.8 + 0.05 * rng.standard_normal((n_rows, n))
    got = bb.upsample_irs(left, right, u, filter=kind)
    h = bb.design_filter(u, kind)
    assert got['upsampling'] == float(u) and got['irs_left'].shape == (n_rows, n * u)
    for ear, x in (('left', left), ('right', right)):
        want = np.stack([bank_oracle.resample(row, u, h) for row in x])
        assert np.abs(got['irs_' + ear] - want).max() <= 1e-12 * np.abs(want).max()
        d = got['diffs_' + ear]
        assert np.array_equal(d, -d.T) and not d.diagonal().any()              # upsample_irs.m:31-32
        for i in range(n_rows):
            for j in range(i + 1, n_rows):
                assert abs(d[i, j] - bank_oracle.delay_difference(x[i], x[j], u, h)) <= 1e-9
    # identical signals have zero delay difference; a pure delay of 3 samples reads as +3
    same = bb.upsample_irs(np.stack([left[0], left[0], np.roll(left[0], 3)]), np.stack([left[0]] * 3), u, filter=kind)
    assert abs(same['diffs_left'][0, 1]) <= 1e-9 and abs(same['diffs_left'][0, 2] - 3.0) <= 1e-3


def test_full_size_bank_matches_the_synthetic_bank_restatement(bas):
    """187 x 512 HRIRs, U = 8, scipy's filter: the device builder reproduces bank_synth.build_bank (the
    numpy/scipy restatement of upsample_irs.m behind every synthetic bank of this repository)."""
    left, right = bas.bank_synth.synthetic_hrirs(0)
    ref = bas.bank_synth.build_bank(8, seed=0)
    got = bas.bank_builder.upsample_irs(left, right, 8, filter='scipy')
    for ear in ('left', 'right'):
        assert np.abs(got['irs_' + ear] - ref['irs_' + ear]).max() <= 1e-12
        assert np.abs(got['diffs_' + ear] - ref['diffs_' + ear]).max() <= 1e-8
    # and a bank built on the device renders like the restated one
    from .conftest import GoldenBank
    keep = 256 * 8
    a = GoldenBank(8, got['diffs_left'], got['diffs_right'], got['irs_left'][:, :keep], got['irs_right'][:, :keep])
    b = GoldenBank(8, ref['diffs_left'], ref['diffs_right'], ref['irs_left'][:, :keep], ref['irs_right'][:, :keep])
    ya = bas.interpolate_2d(a, 0.3, 1.1)
    yb = bas.interpolate_2d(b, 0.3, 1.1)
    assert np.abs(ya - yb).max() <= 1e-6
