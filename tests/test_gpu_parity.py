"""Parity of the CUDA path (through the Python call surface -> C ABI -> sm_100a kernels) against the
oracle and the committed reference vectors.  Tolerances are BASELINE.json's: integer indices
bit-exact; fp32 output within relative L2 error 1e-5 and max-abs error 1e-5 of peak."""
import numpy as np
import pytest

from .conftest import as_kind, golden_trajectory, rel_l2, max_abs_over_peak, KIND_PY, KIND_F64, KIND_F32

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5
MAX_ABS = 1e-5


def close(got, want):
    assert got.shape == want.shape, (got.shape, want.shape)
    assert rel_l2(got, want) <= REL_L2, rel_l2(got, want)
    assert max_abs_over_peak(got, want) <= MAX_ABS, max_abs_over_peak(got, want)


def test_extension_loaded_and_device_present(bas):
    import torch
    assert torch.cuda.is_available()
    assert bas._cabi.lib.bas_device_count() >= 1
    assert torch.cuda.get_device_capability(0)[0] == 10


def test_interpolate_2d_vs_golden(bas, golden, golden_bank):
    for (elev, azim, kind), want in zip(golden['dir_cases'], golden['dir_irs']):
        got = bas.interpolate_2d(golden_bank, elev, as_kind(azim, kind))
        assert got.dtype == np.float64
        close(got, want)


def test_interpolate_2d_batch_device_plan_bit_exact(bas, golden, golden_bank):
    """The fp64 device plan kernel gives the same integers as the reference (and as its host twin)."""
    cases = golden['dir_cases']
    kinds = cases[:, 2].astype(np.uint8)
    filt, trace = bas.interpolate_2d_batch(golden_bank, cases[:, 0], cases[:, 1], kinds, return_trace=True)
    _, host_trace = bas.plan_points_host(golden_bank, cases[:, 0], cases[:, 1], kinds)
    for name in ('rows', 'lo', 'hi', 'err'):
        assert np.array_equal(trace[name], host_trace[name]), name
    for name in ('alpha_top', 'alpha_bot', 'a'):
        assert np.array_equal(trace[name], host_trace[name]), name
    for i, delays in enumerate(golden['dir_delays']):
        for ear in range(2):
            order = [0 + ear, 2 + ear, 4 + ear, 6 + ear, 8 + ear, 10 + ear]
            assert list(trace['lo'][i, ear]) == [int(np.floor(delays[j])) for j in order]
            assert list(trace['hi'][i, ear]) == [int(np.ceil(delays[j])) for j in order]
    close(filt.cpu().numpy().astype(np.float64), golden['dir_irs'])


def test_interpolate_2d_random_vs_oracle(bas, oracle, synth_bank):
    rng = np.random.default_rng(101)
    elev = rng.uniform(-1.2, 1.9, 64)
    azim = rng.uniform(-10, 20, 64)
    filt = bas.interpolate_2d_batch(synth_bank, elev, azim, KIND_F64).cpu().numpy()
    for i in range(64):
        close(filt[i].astype(np.float64), oracle.interpolate_2d(synth_bank, elev[i], np.float64(azim[i])))


def test_interpolate_2d_known_answers(bas, synth_bank):
    """alpha = 0 on a grid ring returns the decimated bank row."""
    u = synth_bank.upsampling
    got = bas.interpolate_2d(synth_bank, 0.0, 0.0)          # row 72 exactly
    want = np.vstack([synth_bank.irs_left[72, ::u], synth_bank.irs_right[72, ::u]])
    close(got, want)
    got = bas.interpolate_2d(synth_bank, np.pi / 2, 1.234)  # the pole: row 186 whatever the azimuth
    want = np.vstack([synth_bank.irs_left[186, ::u], synth_bank.irs_right[186, ::u]])
    close(got, want)


def test_interpolate_2d_errors(bas, golden_bank):
    with pytest.raises(AssertionError):
        bas.interpolate_2d(golden_bank, float('nan'), 1.0)       # apply_hrtf.py:266
    with pytest.raises(AssertionError):
        bas.interpolate_2d(golden_bank, 0.2, float('nan'))       # sphere.py:87
    with pytest.raises(AssertionError):
        bas.interpolate_2d_batch(golden_bank, [0.1, 0.2], [1.0, float('nan')])


def test_ring_interpolation_vs_golden(bas, golden, golden_bank):
    for (b, a, alpha), dec, up, (dl, dr) in zip(golden['ringinterp_in'], golden['ringinterp_dec'],
                                                golden['ringinterp_up'], golden['ringinterp_delays']):
        got = bas.delay_compensated_interpolation_with_delaydiff(golden_bank, int(b), int(a), float(alpha))
        assert got[0] == dl and got[1] == dr            # float64 scalars: exact
        close(got[2], dec)
        got_up = bas.delay_compensated_interpolation_with_delaydiff(golden_bank, int(b), int(a), float(alpha), True)
        close(got_up[2], up)
    with pytest.raises(IndexError):
        bas.delay_compensated_interpolation_with_delaydiff(golden_bank, 187, 0, 0.5)


@pytest.mark.parametrize('variant', ['auto', 'generic'])
def test_make_signal_move_2d_vs_golden(bas, golden, golden_bank, variant):
    k = float(golden['render_k'])
    bas.apply_hrtf.PROGRESS = False
    for i in range(int(golden['n_renders'])):
        n, c, s = (int(v) for v in golden['render%d_meta' % i])
        traj = golden_trajectory(golden['render%d_traj' % i], k)
        x = golden['render%d_x' % i]
        if variant == 'auto':
            got = bas.make_signal_move_2d(x, c, s, traj, golden_bank)
            assert got.shape[1] == 2 and got.dtype == np.float32 and got.flags.f_contiguous
        else:
            got = bas.render_sources(x[None], c, s, [traj], golden_bank, variant=bas._cabi.RENDER_GENERIC)[0].T
        close(got, golden['render%d_y' % i])


def _traj(seed, fs=44100.0):
    rng = np.random.default_rng(seed)
    f1, f2, p1, p2 = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6), rng.uniform(0, 6)
    k = 2 * np.pi / fs

    def fn(t):
        return (np.deg2rad(22.5 + 67.5 * np.sin(f1 * 40 * k * t + p1)), (f2 * 60 * k * t + p2) % (2 * np.pi))
    return fn


_SHAPE_REF = {}


def _shape_reference(oracle, bank, x, traj):
    if 'y' not in _SHAPE_REF:
        _SHAPE_REF['y'] = oracle.make_signal_move_2d(x, 512, 32, traj, bank)
    return _SHAPE_REF['y']


def _shape_reference_n(oracle, bank, x, traj, key):
    if key not in _SHAPE_REF:
        _SHAPE_REF[key] = oracle.make_signal_move_2d(x, 512, 32, traj, bank)
    return _SHAPE_REF[key]


@pytest.mark.parametrize('split', [False, True])
@pytest.mark.parametrize('parts', [1, 2, 4])
@pytest.mark.parametrize('shape', [(4, 2, 2), (4, 1, 2), (4, 1, 3), (6, 1, 2), (6, 2, 1), (8, 2, 1), (8, 1, 1)])
def test_tiled_shapes_vs_oracle(bas, oracle, synth_bank, shape, parts, split):
    """Every compiled tile shape of the register-tiled kernel (warps per CTA x pipeline stages x CTAs
    per SM), with 1, 2 or 4 warps sharing a stripe (tap blocks split between warps, summed in shared
    memory), with whole tiles per CTA and with tiles split between CTAs (stream-K + fix-up), K = 256."""
    tw, ns, ctas = shape
    if tw % parts:
        pytest.skip('parts must divide the warps per CTA')
    rng = np.random.default_rng(5)
    n = 9000
    x = (0.05 * rng.standard_normal(n)).astype(np.float32)
    traj = _traj(3)
    want = _shape_reference(oracle, synth_bank, x, traj)
    variant = bas._cabi.render_variant(tw, ns, ctas, parts, split)
    got = bas.render_sources(x[None], 512, 32, [traj], synth_bank, variant=variant)[0].T
    close(got, want)
    # the same shape mixing three sources (accumulators of the mix live in shared memory)
    xs = np.stack([x, x[::-1].copy(), 0.5 * x])
    trajs = [traj, _traj(4), _traj(5)]
    want_mix = sum(_shape_reference_n(oracle, synth_bank, xs[s], trajs[s], s) for s in range(3))
    try:
        got_mix = bas.render_sources(xs, 512, 32, trajs, synth_bank, mix=True, variant=variant)
    except bas.BasError as e:          # the mix accumulators of this shape do not fit in shared memory
        assert 'no tile shape fits' in str(e)
        return
    close(got_mix.T, want_mix)


@pytest.mark.parametrize('k_taps,c,n', [(100, 512, 5000), (256, 128, 4000), (33, 64, 1500), (512, 512, 6000), (1, 32, 200)])
def test_tiled_odd_geometries_vs_oracle(bas, oracle, k_taps, c, n):
    """samples_to_keep that is not a multiple of 32 (the reference's default is 100), small chunks,
    a one-tap filter."""
    f = bas.bank_synth.build_bank(8, seed=0)
    from .conftest import GoldenBank
    bank = GoldenBank(8, f['diffs_left'], f['diffs_right'], f['irs_left'][:, :k_taps * 8], f['irs_right'][:, :k_taps * 8])
    rng = np.random.default_rng(k_taps)
    x = (0.05 * rng.standard_normal(n)).astype(np.float32)
    traj = _traj(k_taps)
    want = oracle.make_signal_move_2d(x, c, 32, traj, bank)
    got = bas.render_sources(x[None], c, 32, [traj], bank, variant=bas._cabi.RENDER_TILED)[0].T
    close(got, want)
    got = bas.render_sources(x[None], c, 32, [traj], bank, variant=bas._cabi.RENDER_GENERIC)[0].T
    close(got, want)


def test_upsampling_16_bank(bas, oracle):
    """BASELINE.json config 5 shape: N=16 bank, full-length IRs (K = 512)."""
    from .conftest import GoldenBank
    f = bas.bank_synth.build_bank(16, seed=0)
    bank = GoldenBank(16, f['diffs_left'], f['diffs_right'], f['irs_left'], f['irs_right'])
    rng = np.random.default_rng(55)
    x = (0.05 * rng.standard_normal(4000)).astype(np.float32)
    traj = _traj(9)
    close(bas.make_signal_move_2d(x, 512, 32, traj, bank), oracle.make_signal_move_2d(x, 512, 32, traj, bank))


def test_batch_and_mix_vs_oracle(bas, oracle, synth_bank):
    rng = np.random.default_rng(21)
    n_src, n = 5, 6000
    x = (0.02 * rng.standard_normal((n_src, n))).astype(np.float32)
    x[3] *= 400.0                                            # this source peaks above 1 -> normalised on its own
    trajs = [_traj(100 + s) for s in range(n_src)]
    want = [oracle.make_signal_move_2d(x[s], 512, 32, trajs[s], synth_bank).T for s in range(n_src)]
    got = bas.render_sources(x, 512, 32, trajs, synth_bank)
    for s in range(n_src):
        close(got[s], want[s])
    assert abs(np.abs(got[3]).max() - 1.0) < 1e-6
    mix, peaks = bas.render_sources(x, 512, 32, trajs, synth_bank, mix=True, return_peaks=True)
    close(mix, np.sum(np.stack(want).astype(np.float64), axis=0))
    assert peaks[3] > 1 and (np.delete(peaks, 3) < 1).all()
    mix_generic = bas.render_sources(x, 512, 32, trajs, synth_bank, mix=True, variant=bas._cabi.RENDER_GENERIC)
    close(mix_generic, np.sum(np.stack(want).astype(np.float64), axis=0))


def test_time_range_windows(bas, synth_bank):
    rng = np.random.default_rng(31)
    x = (0.05 * rng.standard_normal((1, 7000))).astype(np.float32)
    traj = [_traj(4)]
    for shape in (bas._cabi.render_variant(4, 1, 3, 1), bas._cabi.render_variant(8, 2, 1, 2), bas._cabi.RENDER_GENERIC):
        full = bas.render_sources(x, 512, 32, traj, synth_bank, normalise=False, variant=shape)
        for p0, p1 in [(0, 1), (1, 1000), (777, 4099), (4096, 7423), (7000, 7423)]:
            part = bas.render_sources(x, 512, 32, traj, synth_bank, normalise=False, time_range=(p0, p1), variant=shape)
            assert np.array_equal(part, full[:, :, p0:p1])      # same kernel, same order of operations
    auto = bas.render_sources(x, 512, 32, traj, synth_bank, normalise=False)
    for p0, p1 in [(1, 1000), (4096, 7423)]:     # the library's own choice of tile shape may differ per window
        part = bas.render_sources(x, 512, 32, traj, synth_bank, normalise=False, time_range=(p0, p1))
        assert rel_l2(part, auto[:, :, p0:p1]) <= 1e-6


def test_edge_cases(bas, oracle, golden_bank):
    bas.apply_hrtf.PROGRESS = False
    traj = lambda t: (0, (0.01 * t) % (2 * np.pi))
    # shorter than one chunk, exactly one chunk, impulse (output = first boundary filter)
    for n in (1, 31, 512, 513):
        x = np.zeros(n, dtype=np.float32); x[0] = 1.0
        got = bas.make_signal_move_2d(x, 512, 32, traj, golden_bank)
        want = oracle.make_signal_move_2d(x, 512, 32, traj, golden_bank)
        close(got, want)
        assert got.shape == (int(np.ceil(n / 512)) * 512 + 31, 2)            # apply_hrtf.py:410
    h0 = oracle.interpolate_2d(golden_bank, 0, 0.0)
    x = np.zeros(40, dtype=np.float32); x[0] = 0.5           # peak stays below 1: no normalisation
    got = bas.make_signal_move_2d(x, 512, 32, traj, golden_bank)
    close(got[:32].T.astype(np.float64), 0.5 * h0)
    with pytest.raises(AssertionError):
        bas.make_signal_move_2d(np.zeros((4, 2), dtype=np.float32), 512, 32, traj, golden_bank)   # mono only (:398)
    with pytest.raises(AssertionError):
        bas.make_signal_move_2d(np.zeros(100, dtype=np.float32), 512, 100, traj, golden_bank)    # S | C (:401-402)
    with pytest.raises(AssertionError):
        bas.make_signal_move_2d(np.zeros(1000, dtype=np.float32), 512, 32, lambda t: (0, float('nan')), golden_bank)


def test_vectorised_trajectory_equals_scalar(bas, synth_bank):
    rng = np.random.default_rng(41)
    x = (0.05 * rng.standard_normal(5000)).astype(np.float32)
    k = 2 * np.pi / 2000

    def scalar(t):
        return (np.float64(0.4 * np.sin(k * t)), np.float64((3 * k * t) % (2 * np.pi)))

    def vec(t):
        return (0.4 * np.sin(k * t), (3 * k * t) % (2 * np.pi))
    vec.vectorized = True
    a = bas.render_sources(x[None], 512, 32, [scalar], synth_bank)
    b = bas.render_sources(x[None], 512, 32, [vec], synth_bank)
    assert np.array_equal(a, b)


def test_load_irs_and_delaydiffs_roundtrip(bas, tmp_path, golden_bank):
    """The loader reads the .mat layout upsample_irs.m writes (struct irs_and_delaydiffs, v5)."""
    path = str(tmp_path / 'bank.mat')
    bas.bank_synth.write_mat(path, dict(upsampling=8.0, diffs_left=golden_bank.diffs_left, diffs_right=golden_bank.diffs_right,
                                        irs_left=golden_bank.irs_left, irs_right=golden_bank.irs_right))
    bank = bas.load_irs_and_delaydiffs(path, samples_to_keep=16)
    assert bank.upsampling == 8 and isinstance(bank.upsampling, int)
    assert bank.irs_left.shape == (187, 128) and bank.diffs_right.shape == (187, 187)
    assert np.array_equal(bank.irs_left, golden_bank.irs_left[:, :128])
    got = bas.interpolate_2d(bank, 0.3, 2.0)
    assert got.shape == (2, 16)


@pytest.mark.parametrize('phases,seg_bytes', [((1.0,), 0), ((0.3, 1.0), 1 << 20), ((0.1, 0.35, 0.6, 1.0), 256 << 10)])
def test_host_pipeline_equals_device_path(bas, synth_bank, phases, seg_bytes, monkeypatch):
    """Host arrays in / host array out run as a phased pipeline (bas_pipeline_upload / _phase: upload,
    plan, ir_synth, segmented render and download on three streams).  Cutting the job along time must
    not change a bit: every output sample is computed by the same instructions in the same order as
    in the one-launch, device-resident path."""
    import torch
    ah = bas.apply_hrtf
    monkeypatch.setattr(ah, 'PIPELINE_PHASES', phases)
    monkeypatch.setattr(ah, 'PIPELINE_PHASE_MIN_BYTES', 0)
    monkeypatch.setattr(ah, 'PIPELINE_SEGMENT_BYTES', seg_bytes)
    rng = np.random.default_rng(77)
    n_src, n = 3, 150_000 + 77
    x = (0.02 * rng.standard_normal((n_src, n))).astype(np.float32)
    x[1] *= 300.0                                            # peaks above 1: the second (normalising) pass
    k = 2 * np.pi / 40_000
    def vec(s):
        def fn(t):
            return (np.deg2rad(20 + 60 * np.sin((2 + s) * k * t + s)), ((3 + s) * k * t + 0.5 * s) % (2 * np.pi))
        fn.vectorized = True
        return fn
    trajs = [vec(0), vec(1), lambda t: (np.float64(0.3 * np.cos(k * t)), np.float64((2 * k * t) % (2 * np.pi)))]
    xd = torch.zeros((n_src, (n + 511) // 512 * 512), dtype=torch.float32, device='cuda')
    xd[:, :n] = torch.from_numpy(x).cuda()
    # one tile shape for both paths: the library otherwise picks the shape by the size of each launch,
    # and a different shape may sum the same products in a different order
    shape = bas._cabi.render_variant(4, 1, 2, 1, split=False)
    for mix in (False, True):
        want, want_peaks = bas.render_sources(xd, 512, 32, trajs, synth_bank, mix=mix, return_device=True, return_peaks=True, variant=shape)
        got, got_peaks = bas.render_sources(x, 512, 32, trajs, synth_bank, mix=mix, return_peaks=True, variant=shape)
        assert isinstance(got, np.ndarray) and got.dtype == np.float32
        assert np.array_equal(got, want.cpu().numpy())
        assert np.array_equal(got_peaks, want_peaks) and got_peaks[1] > 1
    # a window of the output, pinned input, and directions given as arrays instead of callables
    times = np.arange(0, xd.shape[1] + 1, 512)
    pre = (np.stack([np.broadcast_to(np.asarray(f(times)[0] if getattr(f, 'vectorized', False) else [f(int(t))[0] for t in times], dtype=np.float64), times.shape) for f in trajs]),
           np.stack([np.asarray(f(times)[1] if getattr(f, 'vectorized', False) else [f(int(t))[1] for t in times], dtype=np.float64) for f in trajs]),
           bas._cabi.AZ_F64)
    xp = torch.from_numpy(x).pin_memory().numpy()
    full = bas.render_sources(xd, 512, 32, pre, synth_bank, normalise=False, return_device=True, variant=shape).cpu().numpy()
    part = bas.render_sources(xp, 512, 32, pre, synth_bank, normalise=False, time_range=(12_345, 140_001), variant=shape)
    assert np.array_equal(part, full[:, :, 12_345:140_001])
    # and with the library's own choice of shapes: same result to rounding
    auto = bas.render_sources(xp, 512, 32, pre, synth_bank, normalise=False)
    assert rel_l2(auto, full) <= 1e-6


def test_legacy_make_signal_move_vs_golden(bas, oracle, golden_bank):
    """The legacy 1-D renderer (apply_hrtf.py:294-353) and delay_compensated_interpolation_easy
    (:114-125) against outputs of the unmodified reference (tests/golden/reference_legacy.npz)."""
    from .test_oracle_golden import _legacy, legacy_index_function
    bas.apply_hrtf.PROGRESS = False
    g = _legacy()
    for v, want in zip(g['easy_in'], g['easy_out']):
        close(bas.delay_compensated_interpolation_easy(golden_bank, float(v)), want)
    for i in range(5):
        name, chunk = g['legacy%d_meta' % i]
        got = bas.make_signal_move(g['legacy%d_x' % i], int(chunk), legacy_index_function(name), golden_bank)
        want = g['legacy%d_y' % i]
        assert got.dtype == np.float32
        close(got, want)
    assert abs(np.abs(g['legacy4_y']).max() - 1.0) < 1e-6          # that case exercises the peak division


def test_cli_writes_the_reference_file_name(bas, oracle, golden_bank, tmp_path, capsys):
    """`python -m binaural_audio_synthesis_b200 in.wav` = the reference's main (apply_hrtf.py:559-649):
    scale by the maximum, fold stereo to mono, render along `passing`, write in-c512-s32-l<K>.wav."""
    from scipy.io import wavfile
    from binaural_audio_synthesis_b200 import cli
    fs = 8000
    rng = np.random.default_rng(9)
    stereo = (3000 * rng.standard_normal((6000, 2))).astype(np.int16)
    wav = tmp_path / 'clip.wav'
    wavfile.write(str(wav), fs, stereo)
    mat = tmp_path / 'bank.mat'
    bas.bank_synth.write_mat(str(mat), dict(upsampling=float(golden_bank.upsampling), diffs_left=golden_bank.diffs_left,
                                            diffs_right=golden_bank.diffs_right, irs_left=golden_bank.irs_left,
                                            irs_right=golden_bank.irs_right))
    assert cli.main([str(wav), '--bank', str(mat), '--samples-to-keep', '32']) == 0
    out_file = tmp_path / 'clip-c512-s32-l32.wav'
    assert out_file.exists() and 'as fast as real time' in capsys.readouterr().out
    fs_out, got = wavfile.read(str(out_file))
    y = stereo.astype(np.float32) / stereo.max()
    mono = 0.5 * y[:, 0] + 0.5 * y[:, 1]
    want = oracle.make_signal_move_2d(mono, 512, 32, cli.trajectories(fs)['passing'], golden_bank)
    assert fs_out == fs and got.dtype == np.float32
    close(got, want)
    assert cli.main([]) == 1                                       # no input file: the reference exits 1 (:570-574)


def test_signal_that_is_not_whole_rows_takes_the_bulk_copy_path(bas, synth_bank):
    """bas_render treats samples >= n_valid as zero.  Whole 32-sample rows arrive by tensor-map TMA
    (hardware zero fill); any other n_valid (a multiple of 4) takes the bulk-copy + re-layout staging.
    Both must give the bits of the explicitly zero-padded signal."""
    import torch
    lib, cabi = bas._cabi.lib, bas._cabi
    dev = bas.apply_hrtf._device_bank(synth_bank)
    rng = np.random.default_rng(99)
    n, c, k = 5000, 512, 256                               # 5000 = 156 rows + 8 samples
    n_in = (n + c - 1) // c * c
    n_out, n_pts = n_in + k - 1, n_in // c + 1
    x = torch.full((n_in,), 7.0, dtype=torch.float32, device='cuda')            # garbage beyond n must not be read
    x[:n] = torch.from_numpy((0.05 * rng.standard_normal(n)).astype(np.float32)).cuda()
    xz = x.clone(); xz[n:] = 0
    times = np.arange(0, n_in + 1, c)
    elev, azim = _traj(6)(times) if getattr(_traj(6), 'vectorized', False) else zip(*[_traj(6)(int(t)) for t in times])
    filt = bas.apply_hrtf._plan_and_synth(torch, dev, torch.tensor(np.asarray(elev, dtype=np.float64)).cuda(),
                                          torch.tensor(np.asarray(azim, dtype=np.float64)).cuda(), cabi.AZ_F64, n_pts, cabi.IR_ROWS)[0]
    stride = (n_out + 3) // 4 * 4
    stream = torch.cuda.current_stream().cuda_stream
    outs = []
    for sig, n_valid in ((x, n), (xz, n_in)):
        out = torch.zeros((2, stride), dtype=torch.float32, device='cuda')
        peak = torch.zeros(1, dtype=torch.float32, device='cuda')
        cabi.check(lib.bas_render(sig.data_ptr(), n_in, n_valid, 1, n_in, c, 32, k, filt.data_ptr(), None, 0, n_out, out.data_ptr(),
                                  stride, 0, peak.data_ptr(), cabi.render_variant(4, 1, 3, 1, False), None, 0, stream), 'bas_render')
        outs.append((out.cpu().numpy(), float(peak)))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert np.abs(outs[0][0]).max() > 0


@pytest.mark.parametrize('seed', range(20))
def test_randomised_geometries_vs_oracle(bas, oracle, seed):
    """Seeded random (K, C, S, N, trajectory) draws: taps that are not multiples of 32 or 4, chunks of a
    single subchunk, subchunk sizes the tiled kernel does not take (generic kernel), signals shorter than
    a chunk, trajectories that wrap the azimuth and leave the elevation range (clamped rings, pole)."""
    rng = np.random.default_rng(1000 + seed)
    k_taps = int(rng.choice([1, 3, 17, 32, 33, 64, 100, 127, 256, 300]))
    s = int(rng.choice([32, 32, 32, 16, 64, 8]))
    c = s * int(rng.choice([1, 2, 4, 16]))
    n = int(rng.integers(1, 6000))
    f = _bank_fields(bas)
    bank = GoldenBankLocal(8, f['diffs_left'], f['diffs_right'], f['irs_left'][:, :k_taps * 8], f['irs_right'][:, :k_taps * 8])
    x = (0.05 * rng.standard_normal(n)).astype(np.float32)
    if seed % 5 == 0:
        x *= 40.0                                               # peak above 1: the division of apply_hrtf.py:462-464
    a0, a1, e0, e1 = rng.uniform(-8, 8), rng.uniform(-0.02, 0.02), rng.uniform(-2, 2), rng.uniform(-0.001, 0.001)
    kind = (float, np.float64, np.float32)[seed % 3]
    traj = lambda t: (e0 + e1 * t, kind((a0 + a1 * t) % (2 * np.pi)))
    want = oracle.make_signal_move_2d(x, c, s, traj, bank)
    got = bas.make_signal_move_2d(x, c, s, traj, bank)
    close(got, want)


_BANK_FIELDS = {}


def _bank_fields(bas):
    if 'f' not in _BANK_FIELDS:
        _BANK_FIELDS['f'] = bas.bank_synth.build_bank(8, seed=0)
    return _BANK_FIELDS['f']


class GoldenBankLocal:
    def __init__(self, upsampling, diffs_left, diffs_right, irs_left, irs_right):
        self.upsampling = int(upsampling)
        self.diffs_left, self.diffs_right = np.asarray(diffs_left, dtype=np.float64), np.asarray(diffs_right, dtype=np.float64)
        self.irs_left, self.irs_right = np.asarray(irs_left, dtype=np.float64), np.asarray(irs_right, dtype=np.float64)
