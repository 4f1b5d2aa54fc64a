"""Round-2 parity cases: delay_signal_float as a callable, the fixes for uninitialised-halo reads,
host tensors, empty input, first-error semantics, and the unmodified-caller form of
make_signal_move_2d (plain lambda + pageable ndarray).  Tolerances are BASELINE.json's."""
import os

import numpy as np
import pytest

from .conftest import rel_l2, max_abs_over_peak, GoldenBank

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5
MAX_ABS = 1e-5


def close(got, want):
    assert got.shape == want.shape, (got.shape, want.shape)
    assert rel_l2(got, want) <= REL_L2, rel_l2(got, want)
    assert max_abs_over_peak(got, want) <= MAX_ABS, max_abs_over_peak(got, want)


def _traj(seed):
    k = 2 * np.pi / 3000
    p = 0.3 + seed
    return lambda t: (np.deg2rad(22.5 + 67.5 * np.sin(3 * k * t + p)), (5 * k * t + 1 + seed) % (2 * np.pi))


def test_delay_signal_float_bit_exact_vs_reference_vectors(bas):
    """apply_hrtf.py:127-165 as its own entry point: float64, circular, bit-identical to vectors made
    by the unmodified reference (negative, integer, beyond-one-period delays; decimation)."""
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_dsf.npz'))
    for i, (n, d, ds) in enumerate(g['cases']):
        got = bas.delay_signal_float(g['x%d' % i], float(d), int(ds))
        assert got.dtype == np.float64
        assert np.array_equal(got, g['y%d' % i]), (i, n, d, ds)
    assert bas.delay_signal_float(np.zeros(0), 1.5).shape == (0,)
    with pytest.raises(ValueError):
        bas.delay_signal_float(np.arange(4.0), float('nan'))            # int(floor(nan)), apply_hrtf.py:149


@pytest.mark.parametrize('k_taps', [100, 37, 256])
def test_time_range_with_poisoned_scratch(bas, oracle, k_taps, monkeypatch):
    """with time_range p0 > 0 and K % 32 != 0 the tiled kernel multiplies input samples
    below p0 - (K - 1) by zero-padding taps; they must be initialised.  All device scratch is filled
    with NaN here, so any such read poisons the output."""
    f = bas.bank_synth.build_bank(8, seed=0)
    bank = GoldenBank(8, f['diffs_left'], f['diffs_right'], f['irs_left'][:, :k_taps * 8], f['irs_right'][:, :k_taps * 8])
    rng = np.random.default_rng(5)
    n = 9000
    x = (0.05 * rng.standard_normal((2, n))).astype(np.float32)
    trajs = [_traj(0), _traj(1)]
    want = np.stack([oracle.make_signal_move_2d(x[s], 512, 32, trajs[s], bank).T for s in range(2)])
    monkeypatch.setattr(bas.apply_hrtf, 'POISON_SCRATCH', True)
    for p0, p1 in [(3000, 6000), (4097, 9100), (8000, want.shape[2])]:
        part = bas.render_sources(x, 512, 32, trajs, bank, normalise=False, time_range=(p0, p1))       # host pipeline
        assert np.isfinite(part).all(), (p0, p1)
        close(part, want[:, :, p0:p1])
        dev = bas.render_sources(x, 512, 32, trajs, bank, normalise=False, time_range=(p0, p1), return_device=True)
        assert np.array_equal(dev.cpu().numpy(), part) or rel_l2(dev.cpu().numpy(), part) <= 1e-6


def test_render_by_time_k100_segments_are_finite(bas, oracle, monkeypatch):
    """distributed.render_by_time's per-rank call (window render with time_range) at the reference's
    default K = 100, scratch poisoned."""
    from binaural_audio_synthesis_b200 import distributed
    f = bas.bank_synth.build_bank(8, seed=0)
    bank = GoldenBank(8, f['diffs_left'], f['diffs_right'], f['irs_left'][:, :800], f['irs_right'][:, :800])
    rng = np.random.default_rng(6)
    n = 512 * 24
    x = (0.05 * rng.standard_normal(n)).astype(np.float32)
    want = oracle.make_signal_move_2d(x, 512, 32, _traj(2), bank).T
    monkeypatch.setattr(bas.apply_hrtf, 'POISON_SCRATCH', True)
    k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
    for p0, p1 in distributed.time_segments(n_in, 512, k, 4):
        n0, n1 = distributed.segment_inputs(p0, p1, n_in, 512, k)
        seg = bas.render_sources(x[None, n0:n1], 512, 32, [distributed._shift_trajectory(_traj(2), n0)], bank, normalise=False,
                                 return_device=True, time_range=(p0 - n0, min(p1, n1 + k - 1) - n0))[0].cpu().numpy()
        assert np.isfinite(seg).all()
        close(seg, want[:, p0:p1])


def test_host_tensors_are_converted_not_reinterpreted(bas, synth_bank):
    """a float64 or non-contiguous CPU tensor must give the same audio as the float32 array."""
    import torch
    rng = np.random.default_rng(8)
    x = (0.05 * rng.standard_normal((2, 5000))).astype(np.float32)
    trajs = [_traj(0), _traj(1)]
    want = bas.render_sources(x, 512, 32, trajs, synth_bank)
    got64 = bas.render_sources(torch.from_numpy(x.astype(np.float64)), 512, 32, trajs, synth_bank)
    assert np.array_equal(got64, want)
    wide = torch.from_numpy(np.repeat(x, 2, axis=1).copy())[:, ::2]            # non-contiguous view of the same samples
    assert not wide.is_contiguous()
    assert np.array_equal(bas.render_sources(wide, 512, 32, trajs, synth_bank), want)


def test_empty_signal_returns_k_minus_one_zero_pairs(bas, golden_bank):
    """The reference pads an empty signal to zero chunks, evaluates traj(0) (apply_hrtf.py:429) and
    returns zeros of length K - 1."""
    bas.apply_hrtf.PROGRESS = False
    k = golden_bank.irs_left.shape[1] // golden_bank.upsampling
    seen = []
    out = bas.make_signal_move_2d(np.zeros(0, dtype=np.float32), 512, 32, lambda t: (seen.append(t), (0.1, 1.0))[1], golden_bank)
    assert out.shape == (k - 1, 2) and out.dtype == np.float32 and not out.any()
    assert seen == [0]
    with pytest.raises(AssertionError):
        bas.make_signal_move_2d(np.zeros(0, dtype=np.float32), 512, 32, lambda t: (0.1, float('nan')), golden_bank)


def test_first_failing_point_decides_the_exception(bas, synth_bank):
    """the reference raises at the FIRST bad trajectory point.  A NaN elevation (vertical
    weight assertion, apply_hrtf.py:266) before a NaN azimuth (sphere.py:87) must report the former,
    and its position."""
    n_pts = 40
    elev = np.full(n_pts, 0.2)
    azim = np.linspace(0, 6, n_pts)
    elev[7] = np.nan                     # -> AssertionError 'interpolation parameter ...' (:266)
    azim[19] = np.nan                    # -> AssertionError 'azim >= 0' (sphere.py:87), later in time
    with pytest.raises(AssertionError, match='interpolation parameter.*direction 7'):
        bas.interpolate_2d_batch(synth_bank, elev, azim)
    elev[7] = 0.2
    with pytest.raises(AssertionError, match='azim >= 0.*direction 19'):
        bas.interpolate_2d_batch(synth_bank, elev, azim)
    x = np.zeros(512 * (n_pts - 1), dtype=np.float32)
    elev[30] = np.nan
    with pytest.raises(AssertionError, match='azim >= 0.*trajectory point 19'):
        bas.render_sources(x[None], 512, 32, (elev[None], azim[None], 1), synth_bank)


def test_unmodified_caller_form_equals_opt_in_form(bas, synth_bank):
    """make_signal_move_2d(pageable ndarray, plain lambda) - what a caller of the reference writes - gives
    bit-identical audio to the pinned + declared-vectorised fast form, and the array is page-locked in
    place only for as long as it lives."""
    import gc
    import torch
    bas.apply_hrtf.PROGRESS = False
    rng = np.random.default_rng(12)
    n = 400_000                                                         # 1.6 MB: above REGISTER_MIN_BYTES
    x = (0.05 * rng.standard_normal(n)).astype(np.float32)
    k = 2 * np.pi / 90000
    plain = lambda t: (np.deg2rad(22.5 + 67.5 * np.sin(3 * k * t + 0.3)), (5 * k * t + 1) % (2 * np.pi))
    fast = lambda t: plain(t)
    fast.vectorized = True
    fast.az_kind = bas._cabi.AZ_PYFLOAT      # `plain` returns Python floats for an int t: float32 ring arithmetic
    before = len(bas.apply_hrtf._registered)
    a = bas.make_signal_move_2d(x, 512, 32, plain, synth_bank)
    assert len(bas.apply_hrtf._registered) == before + 1
    b = bas.make_signal_move_2d(x, 512, 32, plain, synth_bank)          # second call: registration reused
    assert len(bas.apply_hrtf._registered) == before + 1
    x_pinned = torch.from_numpy(x.copy()).pin_memory().numpy()
    c = bas.make_signal_move_2d(x_pinned, 512, 32, fast, synth_bank)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    del x
    gc.collect()
    assert len(bas.apply_hrtf._registered) == before                    # released with the array
    # a float64 signal (converted copy) and a view into a larger array still work
    big = (0.05 * rng.standard_normal(2 * n)).astype(np.float32)
    d = bas.make_signal_move_2d(big[:n], 512, 32, plain, synth_bank)
    e = bas.make_signal_move_2d(big[:n].astype(np.float64), 512, 32, plain, synth_bank)
    assert np.array_equal(d, e)


def test_bank_cache_follows_the_callers_arrays(bas, synth_bank):
    """the device copy of a bank is rebuilt when the caller changes its arrays."""
    bank = GoldenBank(synth_bank.upsampling, synth_bank.diffs_left, synth_bank.diffs_right,
                      synth_bank.irs_left.copy(), synth_bank.irs_right.copy())
    a = bas.interpolate_2d(bank, 0.2, 1.0)
    bank.irs_left *= 2.0
    b = bas.interpolate_2d(bank, 0.2, 1.0)
    assert np.allclose(b[0], 2.0 * a[0], rtol=1e-6) and np.array_equal(b[1], a[1])


@pytest.mark.parametrize('sub', [16, 64, 128, 256])
@pytest.mark.parametrize('fused', [True, False])
def test_tiled_kernel_takes_subchunksize_16_and_multiples_of_32(bas, oracle, synth_bank, sub, fused, monkeypatch):
    """the reference's docstring recommends subchunksize 16 or 32 (apply_hrtf.py:380-381).  The
    tiled kernel (requested explicitly: BAS_RENDER_TILED fails where it does not apply) renders subchunksize 16
    - two blend weights per 32-sample input row - and every multiple of 32 that divides the chunk, fused and not,
    one source per tile and mixing, against the oracle."""
    ah = bas.apply_hrtf
    monkeypatch.setattr(ah, 'FUSED', fused)
    rng = np.random.default_rng(50 + sub)
    n = 9000
    x = (0.05 * rng.standard_normal((3, n))).astype(np.float32)
    trajs = [_traj(s) for s in range(3)]
    want = np.stack([oracle.make_signal_move_2d(x[s], 512, sub, trajs[s], synth_bank).T for s in range(3)])
    for variant in (bas._cabi.RENDER_TILED, bas._cabi.render_variant(8, 2, 1, 1), bas._cabi.render_variant(4, 2, 2, 2)):
        got = bas.render_sources(x, 512, sub, trajs, synth_bank, normalise=False, variant=variant, return_device=True).cpu().numpy()
        close(got, want)
        mix = bas.render_sources(x, 512, sub, trajs, synth_bank, mix=True, normalise=False, variant=variant, return_device=True).cpu().numpy()
        close(mix, want.sum(axis=0))
    # through the drop-in call (host arrays, pipeline): the library's own choice
    bas.apply_hrtf.PROGRESS = False
    got = bas.make_signal_move_2d(x[0], 512, sub, trajs[0], synth_bank)
    close(got.T, oracle.make_signal_move_2d(x[0], 512, sub, trajs[0], synth_bank).T)


def test_subchunksize_8_still_renders_through_the_generic_kernel(bas, oracle, synth_bank):
    rng = np.random.default_rng(58)
    x = (0.05 * rng.standard_normal(3000)).astype(np.float32)
    bas.apply_hrtf.PROGRESS = False
    got = bas.make_signal_move_2d(x, 64, 8, _traj(1), synth_bank)
    close(got, oracle.make_signal_move_2d(x, 64, 8, _traj(1), synth_bank))
    with pytest.raises(bas.BasError):
        bas.render_sources(x[None], 64, 8, [_traj(1)], synth_bank, variant=bas._cabi.RENDER_TILED, return_device=True)


def test_cli_scene_mixes_several_sources(bas, oracle, golden_bank, tmp_path):
    """a scene description (several wav sources, each with its trajectory, gain and delay) rendered
    into one binaural wav = the sum of the reference's make_signal_move_2d outputs of the sources."""
    import json
    import scipy.io
    from scipy.io import wavfile
    from binaural_audio_synthesis_b200 import cli
    fs = 8000
    rng = np.random.default_rng(77)
    a = (0.2 * rng.standard_normal(3000)).astype(np.float32)
    b = (0.2 * rng.standard_normal(2200)).astype(np.float32)
    wavfile.write(str(tmp_path / 'a.wav'), fs, a)
    wavfile.write(str(tmp_path / 'b.wav'), fs, b)
    scipy.io.savemat(str(tmp_path / 'bank.mat'), {'irs_and_delaydiffs': {
        'upsampling': float(golden_bank.upsampling), 'diffs_left': golden_bank.diffs_left, 'diffs_right': golden_bank.diffs_right,
        'irs_left': golden_bank.irs_left, 'irs_right': golden_bank.irs_right}}, format='5', do_compression=False)
    k = golden_bank.irs_left.shape[1] // golden_bank.upsampling
    scene = {'output': str(tmp_path / 'mix.wav'), 'chunksize': 256, 'subchunksize': 32, 'samples_to_keep': k,
             'sources': [{'wav': str(tmp_path / 'a.wav'), 'trajectory': 'circle_horizontal', 'gain': 0.5, 'period': 0.5},
                         {'wav': str(tmp_path / 'b.wav'), 'trajectory': 'passing', 'gain': 1.0, 'delay': 0.05, 'period': 0.5}]}
    (tmp_path / 'scene.json').write_text(json.dumps(scene))
    assert cli.main(['--scene', str(tmp_path / 'scene.json'), '--bank', str(tmp_path / 'bank.mat')]) == 0
    rate, got = wavfile.read(scene['output'])
    assert rate == fs and got.dtype == np.float32
    lead = int(round(0.05 * fs))
    n = max(a.size, b.size + lead)
    xa = np.zeros(n, dtype=np.float32); xa[:a.size] = 0.5 * (a / a.max())
    xb = np.zeros(n, dtype=np.float32); xb[lead:lead + b.size] = b / b.max()
    traj = cli.trajectories(fs, period=0.5)
    want = (oracle.make_signal_move_2d(xa, 256, 32, traj['circle_horizontal'], golden_bank).astype(np.float64) +
            oracle.make_signal_move_2d(xb, 256, 32, lambda t: traj['passing'](t - lead), golden_bank))
    close(got.astype(np.float64), want)


@pytest.mark.parametrize('chunk,sub', [(32, 32), (64, 16), (96, 32), (2048, 64), (4096, 4096)])
def test_chunk_sizes_the_fused_kernel_cannot_stage_fall_back_up_front(bas, oracle, synth_bank, chunk, sub):
    """Tiny chunks mean hundreds of filter rows per tile: no fused shape fits shared memory.  The decision is taken
    before the job starts (bas_render_fused_fits), for host arrays (pipeline.cu) and device tensors alike."""
    import torch
    bas.apply_hrtf.PROGRESS = False
    rng = np.random.default_rng(chunk + sub)
    x = (0.05 * rng.standard_normal(5000)).astype(np.float32)
    want = oracle.make_signal_move_2d(x, chunk, sub, _traj(3), synth_bank)
    close(bas.make_signal_move_2d(x, chunk, sub, _traj(3), synth_bank), want)
    n_in = -(-x.size // chunk) * chunk
    xd = torch.zeros((2, n_in), dtype=torch.float32, device='cuda')
    xd[:, :x.size] = torch.from_numpy(x).cuda()
    mix = bas.render_sources(xd, chunk, sub, [_traj(3), _traj(3)], synth_bank, mix=True, normalise=False, return_device=True).cpu().numpy()
    want_n_in = oracle.make_signal_move_2d(np.concatenate([x, np.zeros(n_in - x.size, dtype=np.float32)]), chunk, sub, _traj(3), synth_bank)
    close(mix.T, 2.0 * want_n_in)


@pytest.mark.gpu
@pytest.mark.parametrize('n_src,seconds,fused', [(8, 20, True), (3, 30, True), (5, 12, False), (48, 3, True), (20, 1, False)])
def test_long_mix_of_few_sources_peaks_and_second_pass(bas, synth_bank, n_src, seconds, fused, monkeypatch):
    """Long mixes with few sources: every tile of the mixing kernel is shared by two CTAs (stream-K spans of 17 - 70
    (tile, source) slices, partial sums through the workspace); short mixes of many sources: fewer tiles than CTAs, the
    sums of a tile travel down a chain of CTAs.  The mix must be the sum of the sources, every per-source peak the
    peak of the complete source (apply_hrtf.py:462), and the second (normalising) pass must use it."""
    import torch
    ah = bas.apply_hrtf
    ah.PROGRESS = False
    monkeypatch.setattr(ah, 'FUSED', 'auto' if fused else False)
    rng = np.random.default_rng(77 + n_src)
    n = seconds * 44100
    x = (0.02 * rng.standard_normal((n_src, n))).astype(np.float32)
    loud = n_src // 2
    x[loud] *= 30.0                                          # peaks above 1: normalised on its own in the second pass
    trajs = [(lambda t, s=s: (0.6 * np.sin(2e-5 * t + s), (3e-5 * t + 0.7 * s) % (2 * np.pi))) for s in range(n_src)]
    xd = torch.from_numpy(x).cuda()
    each, peaks_each = ah.render_sources(xd, 512, 32, trajs, synth_bank, mix=False, normalise=False, return_device=True, return_peaks=True)
    mix, peaks_mix = ah.render_sources(xd, 512, 32, trajs, synth_bank, mix=True, normalise=False, return_device=True, return_peaks=True)
    want = each.double().sum(dim=0)
    assert float((mix.double() - want).norm() / want.norm()) <= 1e-6
    assert np.allclose(peaks_mix, peaks_each, rtol=2e-6, atol=0) and peaks_mix[loud] > 1
    normed = ah.render_sources(xd, 512, 32, trajs, synth_bank, mix=True, return_device=True)
    gains = torch.ones(n_src, dtype=torch.float64, device='cuda')
    gains[loud] = 1.0 / float(peaks_each[loud])
    want_n = (each.double() * gains[:, None, None]).sum(dim=0)
    assert float((normed.double() - want_n).norm() / want_n.norm()) <= 1e-6
    # the same launch twice: identical bits (fixed order of summation, also across the CTA boundary)
    again = ah.render_sources(xd, 512, 32, trajs, synth_bank, mix=True, normalise=False, return_device=True)
    assert torch.equal(again, mix)


@pytest.mark.gpu
def test_render_mix_stream_on_one_rank(bas, synth_bank):
    """distributed.render_mix_stream with a one-rank group: every batch comes out, in order, equal to render_sources'
    mix of the same sources; a failing trajectory raises when its batch is collected.  (The multi-rank exchange is
    checked on real GPUs by tools/dist_check.py and by bench.py's parity field.)"""
    import torch
    import torch.distributed as dist
    ah = bas.apply_hrtf
    ah.PROGRESS = False
    created = not dist.is_initialized()
    if created:
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:29571', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    try:
        rng = np.random.default_rng(5)
        n = 3 * 44100
        batches, want = [], []
        for b in range(4):
            x = torch.from_numpy((0.02 * rng.standard_normal((3, n))).astype(np.float32)).cuda()
            trajs = [(lambda t, s=s, b=b: (0.5 * np.sin(3e-5 * t + s + b), (4e-5 * t + s) % (2 * np.pi))) for s in range(3)]
            batches.append((x, trajs))
            want.append(ah.render_sources(x, 512, 32, trajs, synth_bank, mix=True, normalise=False, return_device=True).clone())
        got = [g.clone() for g in bas.distributed.render_mix_stream(batches, 512, 32, synth_bank)]
        assert len(got) == 4
        for g, w in zip(got, want):
            assert torch.equal(g, w)
        bad = [(batches[0][0], [lambda t: (0.0, float('nan'))] * 3)]
        with pytest.raises(AssertionError):                  # a NaN azimuth fails the reference's `assert azim >= 0` first (sphere.py:87)
            list(bas.distributed.render_mix_stream(bad, 512, 32, synth_bank))
    finally:
        if created:
            dist.destroy_process_group()
