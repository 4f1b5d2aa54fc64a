"""Full-size runs (BASELINE.json config 2 shape: 60 s at 44.1 kHz, K = 256, C = 512, S = 32) checked
through size-independent properties, plus oracle windows: the CPU oracle needs minutes for the
whole signal, so it is run on short windows cut from the long render."""
import numpy as np
import pytest

from .conftest import rel_l2, max_abs_over_peak

pytestmark = pytest.mark.gpu

FS = 44100
N = 60 * FS


def lissajous():
    k = 2 * np.pi / (4 * FS)

    def fn(t):
        return (np.deg2rad(22.5 + 67.5 * np.sin(3 * k * t + 0.3)), (5 * k * t + 1) % (2 * np.pi))
    fn.vectorized = True
    return fn


@pytest.fixture(scope='module')
def long_render(bas, synth_bank):
    rng = np.random.default_rng(2)
    x = (0.05 * rng.standard_normal(N)).astype(np.float32)
    y = bas.render_sources(x[None], 512, 32, [lissajous()], synth_bank, normalise=False)[0]
    return x, y


def test_output_length_and_finiteness(long_render):
    x, y = long_render
    assert y.shape == (2, int(np.ceil(N / 512)) * 512 + 255)
    assert np.isfinite(y).all() and np.abs(y).max() > 0


def test_windows_match_oracle(bas, oracle, synth_bank, long_render):
    """Output window [p0, p1) depends only on inputs [p0-K+1, p1): re-render that stretch with the
    float64 oracle (chunk-aligned, trajectory shifted) and compare."""
    x, y = long_render
    traj = lissajous()
    for p0 in (0, 512 * 777, 512 * 2583, 512 * 5160):
        n0 = max(0, p0 - 512)
        n1 = min(N, p0 + 1536)
        seg = x[n0:n1]
        want = oracle.make_signal_move_2d(seg, 512, 32, lambda t: traj(np.float64(t + n0)), synth_bank).T
        lo = p0 - n0 + (256 if n0 else 0)
        got = y[:, n0 + lo:n0 + seg.size]
        ref = want[:, lo:seg.size]
        assert rel_l2(got, ref) <= 1e-5 and max_abs_over_peak(got, ref) <= 1e-5


def test_linearity_and_tile_shape_independence(bas, synth_bank, long_render):
    x, y = long_render
    traj = [lissajous()]
    y2 = bas.render_sources((2.0 * x)[None], 512, 32, traj, synth_bank, normalise=False,
                            variant=bas._cabi.render_variant(tw=6, parts=2))[0]
    # scaling by 2 is exact in binary floating point; a different tile shape only changes the
    # order of partial sums
    assert rel_l2(y2, 2.0 * y) <= 1e-6


def test_silence_and_impulse_train(bas, oracle, synth_bank):
    traj = lissajous()
    x = np.zeros(N, dtype=np.float32)
    assert not bas.render_sources(x[None], 512, 32, [traj], synth_bank).any()
    # impulses K apart: each output stretch is exactly the filter of the impulse's subchunk
    pos = np.arange(100, N - 600, 300007)
    x[pos] = 1.0
    y = bas.render_sources(x[None], 512, 32, [traj], synth_bank, normalise=False)[0]
    for p in pos[:6]:
        c = p // 512
        h0 = oracle.interpolate_2d(synth_bank, *traj(np.float64(c * 512)))
        h1 = oracle.interpolate_2d(synth_bank, *traj(np.float64((c + 1) * 512)))
        alpha = ((p % 512) // 32 * 32) / 512
        want = (1 - alpha) * h0 + alpha * h1
        got = y[:, p:p + 256]
        assert rel_l2(got, want) <= 1e-5 and max_abs_over_peak(got, want) <= 1e-5
