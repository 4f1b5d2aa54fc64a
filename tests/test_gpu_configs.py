"""BASELINE.json configs 3, 4 and 5 as parity cases on one GPU, at sizes the box renders in a moment,
checked through properties that do not depend on the size (mix = sum of the parts, a time-sharded
render stitches to the whole, seams and random windows against the float64 oracle)."""
import numpy as np
import pytest

from .conftest import GoldenBank, rel_l2, max_abs_over_peak

pytestmark = pytest.mark.gpu


def _lissajous(fs, seed):
    rng = np.random.default_rng(1000 + seed)
    f1, f2, p1, p2 = rng.uniform(1, 4), rng.uniform(2, 6), rng.uniform(0, 6), rng.uniform(0, 6)
    k = 2 * np.pi / (4 * fs)

    def fn(t):
        return (np.deg2rad(22.5 + 67.5 * np.sin(f1 * k * t + p1)), (f2 * k * t + p2) % (2 * np.pi))
    fn.vectorized = True
    return fn


def _oracle_window(oracle, bank, x, traj, p0, width, chunk=512):
    """Oracle output samples [p0, p0 + width) of the render of x: re-render the inputs the window
    depends on (chunk-aligned, trajectory shifted), K - 1 samples of history first."""
    k = bank.irs_left.shape[1] // bank.upsampling
    n0 = max(0, (p0 - k) // chunk * chunk)
    n1 = min(x.size, p0 + width)
    seg = x[n0:n1]
    y = oracle.make_signal_move_2d(seg, chunk, 32, lambda t: traj(np.float64(t + n0)), bank).T     # (2, ...)
    return y[:, p0 - n0:p0 - n0 + width]


def test_config3_many_sources_mixed(bas, oracle, synth_bank):
    """64 independent sources, each on its own trajectory, mixed to one binaural output."""
    fs, n_src, n = 44100, 64, 3 * 44100
    rng = np.random.default_rng(3)
    x = (0.05 / 8 * rng.standard_normal((n_src, n))).astype(np.float32)
    trajs = [_lissajous(fs, s) for s in range(n_src)]
    mix = bas.render_sources(x, 512, 32, trajs, synth_bank, mix=True)
    each = bas.render_sources(x, 512, 32, trajs, synth_bank, mix=False)
    assert mix.shape == (2, each.shape[2]) and each.shape[0] == n_src
    assert rel_l2(mix, each.astype(np.float64).sum(axis=0)) <= 1e-6            # the mix is the sum of its parts
    # source-sharded partial mixes (what every rank contributes to the NCCL sum) add up to the same
    parts = [bas.render_sources(x[r::4], 512, 32, trajs[r::4], synth_bank, mix=True).astype(np.float64) for r in range(4)]
    assert rel_l2(sum(parts), mix) <= 1e-6
    # a window of the mix against the oracle
    p0, width = 512 * 100 + 17, 700
    want = sum(_oracle_window(oracle, synth_bank, x[s], trajs[s], p0, width) for s in range(n_src))
    got = mix[:, p0:p0 + width]
    assert rel_l2(got, want) <= 1e-5 and max_abs_over_peak(got, want) <= 1e-5


def test_config4_long_source_cut_in_time(bas, oracle, synth_bank):
    """One long 48 kHz source cut into 8 time segments with an IR-length input halo: every segment is
    rendered from its own window of the signal and the pieces stitch to the one-piece render."""
    fs, n = 48000, 100 * 48000 + 321
    rng = np.random.default_rng(4)
    x = (0.05 * rng.standard_normal(n)).astype(np.float32)
    k = 2 * np.pi / (60 * fs)
    spiral = lambda t: (np.deg2rad(-40.0) + (np.deg2rad(125.0) / n) * np.asarray(t, dtype=np.float64), (40 * k * np.asarray(t, dtype=np.float64)) % (2 * np.pi))
    spiral.vectorized = True
    shape = bas._cabi.render_variant(4, 1, 3, 1, split=False)
    whole = bas.render_sources(x[None], 512, 32, [spiral], synth_bank, normalise=False, variant=shape)[0]
    taps, n_in, n_out = bas.render_geometry(n, 512, 32, synth_bank)
    assert whole.shape == (2, n_out)
    dist = bas.distributed
    stitched = np.zeros_like(whole)
    for p0, p1 in dist.time_segments(n_in, 512, taps, 8):
        n0, n1 = dist.segment_inputs(p0, p1, n_in, 512, taps)
        window = np.zeros(n1 - n0, dtype=np.float32)
        window[:min(n1, n) - n0] = x[n0:min(n1, n)]
        seg = bas.render_sources(window[None], 512, 32, [dist._shift_trajectory(spiral, n0)], synth_bank, normalise=False,
                                 time_range=(p0 - n0, min(p1, n1 + taps - 1) - n0), variant=shape)[0]
        stitched[:, p0:p0 + seg.shape[1]] = seg
    assert np.array_equal(stitched, whole)                 # same products in the same order on both sides of every cut
    # oracle at the seams and at the very end
    cuts = [p0 for p0, _ in dist.time_segments(n_in, 512, taps, 8)][1:]
    for p in cuts[:3] + [n_out - 600]:
        want = _oracle_window(oracle, synth_bank, x, spiral, p - 100 if p in cuts else p, 600)
        got = whole[:, (p - 100 if p in cuts else p):(p - 100 if p in cuts else p) + 600]
        assert rel_l2(got, want) <= 1e-5 and max_abs_over_peak(got, want) <= 1e-5


def test_config5_full_length_irs_n16_bank(bas, oracle):
    """N = 16 upsampled bank with full-length IRs (samples_to_keep = 512, K = 512), many sources."""
    f = bas.bank_synth.build_bank(16, seed=0)
    bank = GoldenBank(16, f['diffs_left'], f['diffs_right'], f['irs_left'], f['irs_right'])
    fs, n_src, n = 44100, 24, 2 * 44100 + 99
    rng = np.random.default_rng(5)
    x = (0.01 * rng.standard_normal((n_src, n))).astype(np.float32)
    trajs = [_lissajous(fs, 50 + s) for s in range(n_src)]
    each = bas.render_sources(x, 512, 32, trajs, bank)
    assert each.shape == (n_src, 2, (n + 511) // 512 * 512 + 511)
    mix = bas.render_sources(x, 512, 32, trajs, bank, mix=True)
    assert rel_l2(mix, each.astype(np.float64).sum(axis=0)) <= 1e-6
    for s in (0, 11, 23):                                   # a random subset of sources against the oracle
        p0 = 512 * (20 + s) + 5
        want = _oracle_window(oracle, bank, x[s], trajs[s], p0, 900)
        got = each[s][:, p0:p0 + 900]
        assert rel_l2(got, want) <= 1e-5 and max_abs_over_peak(got, want) <= 1e-5
    # linearity in the signal: exact for a power of two when the tile shape (summation order) is the same
    shape = bas._cabi.render_variant(4, 1, 2, 1, split=False)
    once = bas.render_sources(x[:3], 512, 32, trajs[:3], bank, variant=shape)
    twice = bas.render_sources(2.0 * x[:3], 512, 32, trajs[:3], bank, variant=shape)
    assert np.array_equal(twice, 2.0 * once)
    assert rel_l2(once, each[:3]) <= 1e-6


def test_fused_filter_synthesis_is_bit_identical(bas, synth_bank, monkeypatch):
    """The render kernel that synthesises its own filter rows (bas_render_fused, the default) against the
    two-kernel path (bas_ir_synth writes the rows to HBM, bas_render copies them): same terms, same order of
    summation, so the audio is bit-identical for a given tile shape - mixing and not, with a source that is
    normalised on its own (gains pass), and a failing direction is reported with its source."""
    import torch
    ah = bas.apply_hrtf
    fs, n_src, n = 44100, 37, 40_000
    rng = np.random.default_rng(8)
    x = (0.01 * rng.standard_normal((n_src, n))).astype(np.float32)
    x[20] *= 300.0                                             # one source that is normalised on its own (gains pass)
    trajs = [_lissajous(fs, 200 + s) for s in range(n_src)]
    xd = torch.zeros((n_src, (n + 511) // 512 * 512), dtype=torch.float32, device='cuda')
    xd[:, :n] = torch.from_numpy(x).cuda()
    for mix in (True, False):
        for shape in (bas._cabi.render_variant(4, 2, 2, 1), bas._cabi.render_variant(8, 2, 1, 2), bas._cabi.render_variant(4, 2, 2, 1, split=True), bas._cabi.render_variant(8, 2, 1, 1)):
            monkeypatch.setattr(ah, 'FUSED', False)
            two, peaks_two = ah.render_sources(xd, 512, 32, trajs, synth_bank, mix=mix, return_device=True, return_peaks=True, variant=shape)
            monkeypatch.setattr(ah, 'FUSED', True)
            one, peaks_one = ah.render_sources(xd, 512, 32, trajs, synth_bank, mix=mix, return_device=True, return_peaks=True, variant=shape)
            assert np.array_equal(peaks_one, peaks_two) and peaks_one[20] > 1
            assert bas._cabi.lib.bas_render_fused_shape(shape) == 1     # the comparison is not the two-kernel path against itself
            assert torch.equal(one, two), (mix, shape)
    # host arrays (pipeline.cu) take the fused kernel too
    monkeypatch.setattr(ah, 'FUSED', False)
    two = ah.render_sources(x[:5], 512, 32, trajs[:5], synth_bank, mix=True)
    monkeypatch.setattr(ah, 'FUSED', True)
    one = ah.render_sources(x[:5], 512, 32, trajs[:5], synth_bank, mix=True)
    assert rel_l2(one, two) <= 1e-6                             # the library picks the tile shape per call
    bad = list(trajs)
    bad[30] = lambda t: (0.0, float('nan'))
    with pytest.raises(AssertionError, match='source 30'):
        ah.render_sources(xd, 512, 32, bad, synth_bank, mix=True, return_device=True)
