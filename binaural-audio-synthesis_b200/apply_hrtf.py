"""Host side of the B200 moving-source binaural renderer: the call surface of the reference's
apply_hrtf.py for its hot path, on hand-written sm_100a kernels behind a C ABI.

Drop-in functions (same names, arguments, return types and error behaviour as the reference):

    load_irs_and_delaydiffs(filename, samples_to_keep)                          apply_hrtf.py:23-46
    delay_compensated_interpolation_with_delaydiff(bank, before, after, alpha,
                                                   return_upsampled=False)       apply_hrtf.py:53-106
    interpolate_2d(bank, elev, azim)                                            apply_hrtf.py:171-281
    make_signal_move_2d(in_signal, chunksize, subchunksize, elev_azim_function,
                        bank)                                                    apply_hrtf.py:356-466

plus thin wrappers the reference also has (delay_compensated_interpolation, ..._easy,
interpolate_2d_deg) and batched entry points that the reference lacks but a GPU needs
(interpolate_2d_batch, render_sources).

PyTorch supplies device memory and streams only; all arithmetic is in libbas_b200.so.  There is no
CPU path: without a CUDA device every compute call raises.
"""
from __future__ import annotations

import math

import numpy as np

from . import _cabi
from . import sphere
from ._cabi import BasError, lib

PROGRESS = True          # print the reference's closing ' 100.0%' line (apply_hrtf.py:457)


# --------------------------------------------------------------------------------------------------
# bank
# --------------------------------------------------------------------------------------------------
class _BankDevice:
    """HBM-resident form of a bank: polyphase fp32 HRIRs for both ears and the fp64 delay tables."""

    def __init__(self, bank, torch, device):
        u = int(bank.upsampling)
        irs_l = np.ascontiguousarray(bank.irs_left, dtype=np.float64)
        irs_r = np.ascontiguousarray(bank.irs_right, dtype=np.float64)
        if irs_l.shape != irs_r.shape or irs_l.shape[0] != _cabi.N_DIRECTIONS:
            raise ValueError('bank must hold %d directions per ear' % _cabi.N_DIRECTIONS)
        length = irs_l.shape[1]
        if length % u:
            raise ValueError('row length %d is not a multiple of the upsampling factor %d' % (length, u))
        self.upsampling, self.length, self.taps = u, length, length // u
        self.device = device
        self.diffs_host = np.ascontiguousarray(
            np.stack([np.asarray(bank.diffs_left, dtype=np.float64), np.asarray(bank.diffs_right, dtype=np.float64)]))
        if self.diffs_host.shape != (2, _cabi.N_DIRECTIONS, _cabi.N_DIRECTIONS):
            raise ValueError('delay-difference tables must be 187 x 187')
        stream = torch.cuda.current_stream(device).cuda_stream
        self.diffs = torch.from_numpy(self.diffs_host).to(device)
        self.bank_pp = torch.empty((2, _cabi.N_DIRECTIONS, u, self.taps), dtype=torch.float32, device=device)
        for ear, irs in enumerate((irs_l, irs_r)):
            staged = torch.from_numpy(irs).to(device)
            _cabi.check(lib.bas_bank_to_polyphase(staged.data_ptr(), _cabi.N_DIRECTIONS, length, u,
                                                  self.bank_pp[ear].data_ptr(), stream), 'bas_bank_to_polyphase')
        torch.cuda.current_stream(device).synchronize()     # staged fp64 rows may now be freed


def _device_bank(bank) -> _BankDevice:
    """Device state of `bank`, uploaded on first use and cached on the object (one per device).
    Works for the object load_irs_and_delaydiffs returns and for any object carrying the
    reference's five attributes (apply_hrtf.py:36-44)."""
    torch = _cabi.require_device()
    device = torch.device('cuda', torch.cuda.current_device())
    cache = bank.__dict__.get('_bas_device') if hasattr(bank, '__dict__') else None
    if cache is None:
        cache = {}
        try:
            setattr(bank, '_bas_device', cache)
        except (AttributeError, TypeError):
            pass
    state = cache.get(device.index)
    if state is None:
        state = _BankDevice(bank, torch, device)
        cache[device.index] = state
    return state


def load_irs_and_delaydiffs(filename='irs_and_delaydiffs_compensated_6.mat', samples_to_keep=512):
    """apply_hrtf.py:23-46.  Reads the MATLAB-v5 struct upsample_irs.m writes, keeps the first
    samples_to_keep*upsampling samples of every row, and returns a class object with the five
    attributes of the reference (upsampling, diffs_left, diffs_right, irs_left, irs_right).  When a
    CUDA device is present the bank is uploaded to HBM immediately (once); otherwise the upload
    happens on first use."""
    import scipy.io
    m = scipy.io.loadmat(filename)['irs_and_delaydiffs']

    class irs_and_delaydiffs:
        upsampling = int(m[0][0]['upsampling'][0][0])
        diffs_left = m[0][0]['diffs_left']
        diffs_right = m[0][0]['diffs_right']
        irs_left = m[0][0]['irs_left'][:, :samples_to_keep * upsampling]
        irs_right = m[0][0]['irs_right'][:, :samples_to_keep * upsampling]

    try:
        import torch
        if torch.cuda.is_available():
            _device_bank(irs_and_delaydiffs)
    except ImportError:
        pass
    return irs_and_delaydiffs


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _stream(torch):
    return torch.cuda.current_stream().cuda_stream


def _raise_plan_error(err: int, where: str = ''):
    if err & _cabi.ERR_AZIM_ASSERT:
        raise AssertionError('azim >= 0' + where)                                      # sphere.py:87
    if err & _cabi.ERR_VERT_ASSERT:
        raise AssertionError('interpolation parameter somehow takes invalid value' + where)   # apply_hrtf.py:266
    if err & _cabi.ERR_NONFINITE:
        raise ValueError('cannot convert float NaN to integer' + where)                # apply_hrtf.py:149
    if err:
        raise BasError('plan error %d%s' % (err, where))


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def plan_points_host(bank, elev, azim, kinds):
    """Host twin of the device plan kernel (same plan_math.h): returns (terms, trace) numpy
    structured arrays for the given points.  Needs no GPU; used by the scalar entry points and by
    index-parity tests."""
    elev = np.ascontiguousarray(elev, dtype=np.float64).ravel()
    azim = np.ascontiguousarray(azim, dtype=np.float64).ravel()
    n = elev.size
    if np.isscalar(kinds):
        kinds_arr, kind_all = None, int(kinds)
    else:
        kinds_arr, kind_all = np.ascontiguousarray(kinds, dtype=np.uint8).ravel(), 0
    diffs = np.ascontiguousarray(np.stack([np.asarray(bank.diffs_left, dtype=np.float64),
                                           np.asarray(bank.diffs_right, dtype=np.float64)]))
    u = int(bank.upsampling)
    length = int(bank.irs_left.shape[1])
    terms = np.zeros((n, 2, _cabi.MAX_TERMS), dtype=_cabi.TERM_DTYPE)
    trace = np.zeros(n, dtype=_cabi.TRACE_DTYPE)
    rc = lib.bas_plan_build_host(diffs[0].ctypes.data, diffs[1].ctypes.data, u, length,
                                 elev.ctypes.data, azim.ctypes.data,
                                 kinds_arr.ctypes.data if kinds_arr is not None else None, kind_all, n,
                                 terms.ctypes.data, trace.ctypes.data)
    if rc < 0:
        raise BasError(_cabi.last_error())
    return terms, trace


# --------------------------------------------------------------------------------------------------
# ring interpolation (apply_hrtf.py:53-125)
# --------------------------------------------------------------------------------------------------
def _grid_row(index) -> int:
    i = int(index)
    if i < 0:
        i += _cabi.N_DIRECTIONS                      # numpy-style negative index
    if not 0 <= i < _cabi.N_DIRECTIONS:
        raise IndexError('index %d is out of bounds for axis 0 with size %d' % (int(index), _cabi.N_DIRECTIONS))
    return i


def delay_compensated_interpolation_with_delaydiff(irs_and_delaydiffs, before: int, after: int, alpha: float,
                                                   return_upsampled=False):
    """apply_hrtf.py:53-106: delay-compensated interpolation between grid rows `before` and `after`
    with weight `alpha`.  Returns (delay_l / U, delay_r / U, irs) with irs of shape (2, K), or
    (2, K*U) when return_upsampled."""
    torch = _cabi.require_device()
    dev = _device_bank(irs_and_delaydiffs)
    b, a = _grid_row(before), _grid_row(after)
    # (1 - alpha) is evaluated in alpha's own precision by the reference (apply_hrtf.py:90-91)
    one_minus = float(1 - alpha)
    terms = np.zeros((1, 2, _cabi.MAX_TERMS), dtype=_cabi.TERM_DTYPE)
    delays = np.zeros(2, dtype=np.float64)
    rc = lib.bas_plan_ring_host(dev.diffs_host[0].ctypes.data, dev.diffs_host[1].ctypes.data, dev.upsampling,
                                dev.length, b, a, float(alpha), one_minus, terms.ctypes.data, delays.ctypes.data,
                                None, None)
    if rc < 0:
        raise BasError(_cabi.last_error())
    _raise_plan_error(rc)
    width = dev.length if return_upsampled else dev.taps
    terms_dev = torch.from_numpy(terms.view(np.uint8).reshape(-1)).to(dev.device)
    out = torch.empty((1, 2, width), dtype=torch.float32, device=dev.device)
    _cabi.check(lib.bas_ir_synth(dev.bank_pp.data_ptr(), dev.upsampling, dev.taps, terms_dev.data_ptr(), 1,
                                 _cabi.IR_UPSAMPLED if return_upsampled else _cabi.IR_PLANAR, out.data_ptr(), width,
                                 _stream(torch)), 'bas_ir_synth')
    irs = out[0].cpu().numpy().astype(np.float64)
    return (delays[0], delays[1], irs)


def delay_compensated_interpolation(irs_and_delaydiffs, before: int, after: int, alpha: float):
    """apply_hrtf.py:108-111."""
    return delay_compensated_interpolation_with_delaydiff(irs_and_delaydiffs, before, after, alpha)[2]


def delay_compensated_interpolation_easy(irs_and_delaydiffs, continuous_index: float):
    """apply_hrtf.py:114-125 (including its 97 -> 73 wrap of the horizontal ring)."""
    before = int(np.floor(continuous_index))
    after = int(np.ceil(continuous_index))
    alpha = continuous_index - before
    if after == 97:
        after = 73
    return delay_compensated_interpolation(irs_and_delaydiffs, before, after, alpha)


# --------------------------------------------------------------------------------------------------
# 2-D interpolation (apply_hrtf.py:167-281)
# --------------------------------------------------------------------------------------------------
def interpolate_2d_batch(irs_and_delaydiffs, elev, azim, az_kind=_cabi.AZ_F64, return_trace=False, check=True):
    """interpolate_2d for arrays of directions, entirely on the device.

    elev, azim: array-likes of n radians values (numpy or CUDA torch float64).  az_kind: one
    _cabi.AZ_* value or an array of n of them (which arithmetic sphere.py would have used for each
    azimuth object, see sphere.az_kind).  Returns a CUDA float32 tensor (n, 2, K); with
    return_trace also the per-point integer trace (numpy structured array)."""
    torch = _cabi.require_device()
    dev = _device_bank(irs_and_delaydiffs)
    elev_d = torch.as_tensor(elev, dtype=torch.float64).reshape(-1).to(dev.device).contiguous()
    azim_d = torch.as_tensor(azim, dtype=torch.float64).reshape(-1).to(dev.device).contiguous()
    n = elev_d.numel()
    if azim_d.numel() != n:
        raise ValueError('elev and azim must have the same number of points')
    filt, status, trace = _plan_and_synth(torch, dev, elev_d, azim_d, az_kind, n, _cabi.IR_PLANAR, return_trace)
    if check:
        err, where = (int(v) for v in status.cpu())
        if err:
            _raise_plan_error(err, ' (direction %d)' % where)
    if return_trace:
        return filt, trace.cpu().numpy().view(_cabi.TRACE_DTYPE).reshape(n)
    return filt


def _plan_and_synth(torch, dev, elev_d, azim_d, az_kind, n, mode, want_trace=False):
    """plan_build + ir_synth for n directions already on the device; returns (filters, status
    int32[2], trace bytes or None), all asynchronous.  mode IR_PLANAR: filters (n, 2, K);
    IR_ROWS: (n, pitch, 2) filter rows for bas_render."""
    stream = _stream(torch)
    if np.isscalar(az_kind):
        kinds_ptr, kind_all, kinds_d = None, int(az_kind), None
    else:
        kinds_d = torch.as_tensor(np.ascontiguousarray(az_kind, dtype=np.uint8).reshape(-1)).to(dev.device)
        if kinds_d.numel() != n:
            raise ValueError('az_kind must have one entry per direction')
        kinds_ptr, kind_all = kinds_d.data_ptr(), 0
    terms = torch.empty((max(n, 1), 2 * _cabi.MAX_TERMS * 8), dtype=torch.uint8, device=dev.device)
    status = torch.empty(2, dtype=torch.int32, device=dev.device)
    trace = torch.empty((max(n, 1), np.dtype(_cabi.TRACE_DTYPE).itemsize), dtype=torch.uint8,
                        device=dev.device) if want_trace else None
    _cabi.check(lib.bas_plan_build(dev.diffs[0].data_ptr(), dev.diffs[1].data_ptr(), dev.upsampling, dev.length,
                                   elev_d.data_ptr(), azim_d.data_ptr(), kinds_ptr, kind_all, n, terms.data_ptr(),
                                   trace.data_ptr() if want_trace else None, status.data_ptr(), stream),
                'bas_plan_build')
    if mode == _cabi.IR_ROWS:
        filt = torch.empty((n, lib.bas_filter_row_pitch(dev.taps), 2), dtype=torch.float32, device=dev.device)
    else:
        filt = torch.empty((n, 2, dev.taps), dtype=torch.float32, device=dev.device)
    _cabi.check(lib.bas_ir_synth(dev.bank_pp.data_ptr(), dev.upsampling, dev.taps, terms.data_ptr(), n, mode,
                                 filt.data_ptr(), dev.taps, stream), 'bas_ir_synth')
    return filt, status, trace


def interpolate_2d(irs_and_delaydiffs, elev, azim):
    """apply_hrtf.py:171-281: the HRIR pair for direction (elev, azim) in radians, shape (2, K)."""
    torch = _cabi.require_device()
    dev = _device_bank(irs_and_delaydiffs)
    kind = sphere.az_kind(azim)
    terms, trace = plan_points_host(irs_and_delaydiffs, [float(elev)], [float(azim)], kind)
    _raise_plan_error(int(trace['err'][0]))
    terms_dev = torch.from_numpy(terms.view(np.uint8).reshape(-1)).to(dev.device)
    out = torch.empty((1, 2, dev.taps), dtype=torch.float32, device=dev.device)
    _cabi.check(lib.bas_ir_synth(dev.bank_pp.data_ptr(), dev.upsampling, dev.taps, terms_dev.data_ptr(), 1,
                                 _cabi.IR_PLANAR, out.data_ptr(), dev.taps, _stream(torch)), 'bas_ir_synth')
    return out[0].cpu().numpy().astype(np.float64)


def interpolate_2d_deg(irs_and_delaydiffs, elev, azim):
    """apply_hrtf.py:167-169."""
    deg2rad = 2 * np.pi / 360
    return interpolate_2d(irs_and_delaydiffs, elev * deg2rad, azim * deg2rad)


# --------------------------------------------------------------------------------------------------
# renderer (apply_hrtf.py:356-466)
# --------------------------------------------------------------------------------------------------
def render_geometry(n_samples: int, chunksize: int, subchunksize: int, irs_and_delaydiffs):
    """(K, N_in, N_out) with the reference's own expressions and assertions (apply_hrtf.py:399-411)."""
    ir_length = int(0.5 + irs_and_delaydiffs.irs_left.shape[1] / irs_and_delaydiffs.upsampling)
    chunks_per_subchunk = chunksize / subchunksize
    assert chunks_per_subchunk == np.floor(chunks_per_subchunk), 'subchunksize does not divide chunksize evenly'
    in_length = int(0.5 + np.ceil(n_samples / chunksize) * chunksize)
    out_length = in_length + ir_length - 1
    return ir_length, in_length, out_length


def evaluate_trajectory(elev_azim_function, times):
    """Directions at the chunk boundaries.  The reference calls elev_azim_function(t) with a Python
    int for t = 0, C, ..., N_in (apply_hrtf.py:429, :435).  A callable that sets
    `vectorized = True` is called once with the whole int64 array instead and must return two
    arrays; its azimuths are treated as float64 scalars unless it also sets `az_kind`.
    Returns (elev float64[n], azim float64[n], kinds uint8[n] or a single kind)."""
    if getattr(elev_azim_function, 'vectorized', False):
        elev, azim = elev_azim_function(np.asarray(times, dtype=np.int64))
        azim = np.asarray(azim)
        kind = getattr(elev_azim_function, 'az_kind', None)
        if kind is None:
            kind = _cabi.AZ_F32 if azim.dtype == np.float32 else _cabi.AZ_F64
        elev = np.broadcast_to(np.asarray(elev, dtype=np.float64), (len(times),))
        azim = np.broadcast_to(azim.astype(np.float64), (len(times),))
        return np.ascontiguousarray(elev), np.ascontiguousarray(azim), int(kind)
    n = len(times)
    elev = np.empty(n, dtype=np.float64)
    azim = np.empty(n, dtype=np.float64)
    kinds = np.empty(n, dtype=np.uint8)
    for i, t in enumerate(times):
        e, a = elev_azim_function(int(t))            # a Python int, like range() gives the reference
        elev[i] = e
        azim[i] = a
        kinds[i] = sphere.az_kind(a)
    if n and (kinds == kinds[0]).all():
        return elev, azim, int(kinds[0])
    return elev, azim, kinds


def render_sources(signals, chunksize: int, subchunksize: int, elev_azim_functions, irs_and_delaydiffs,
                   mix=False, normalise=True, variant=_cabi.RENDER_AUTO, return_device=False,
                   time_range=None, return_peaks=False):
    """Render n_src mono signals of equal length, each along its own trajectory, in one batch.

    signals: (n_src, N) array-like (numpy, or a CUDA float32 torch tensor for device-resident input).
    elev_azim_functions: one callable per source (see evaluate_trajectory), or a tuple
        (elev, azim, az_kind) of pre-evaluated (n_src, N_in/C + 1) arrays (numpy or CUDA float64).
    mix=False: returns (n_src, 2, N_out) float32, every source normalised like the reference
        (apply_hrtf.py:462-464) when `normalise`.
    mix=True: returns (2, N_out): the sum over sources of what make_signal_move_2d returns for each
        (sources whose own peak exceeds 1 enter the sum divided by that peak when `normalise`).
    time_range=(p0, p1): only output samples p0 <= p < p1 are rendered and returned.
    return_device: return CUDA tensors instead of numpy arrays (no host copy)."""
    torch = _cabi.require_device()
    dev = _device_bank(irs_and_delaydiffs)
    device = dev.device
    # ---- signals -> HBM, zero padded to a multiple of chunksize (apply_hrtf.py:405-406) ----------
    if isinstance(signals, torch.Tensor):
        src = signals
    else:
        src = torch.from_numpy(np.ascontiguousarray(signals, dtype=np.float32))
    if src.dim() != 2:
        raise ValueError('signals must be (n_src, N)')
    n_src, n = src.shape
    k, n_in, n_out = render_geometry(n, chunksize, subchunksize, irs_and_delaydiffs)
    assert k == dev.taps
    if src.is_cuda and src.dtype == torch.float32 and src.is_contiguous() and n == n_in and src.device == device:
        x = src
    else:
        x = torch.empty((n_src, n_in), dtype=torch.float32, device=device)
        if n_in > n:
            x[:, n:].zero_()
        # pinned host memory is copied by DMA without a staging pass; pageable memory goes through
        # the driver's bounce buffers
        x[:, :n].copy_(src, non_blocking=(not src.is_cuda) and src.is_pinned())
    n_pts = n_in // chunksize + 1
    times = np.arange(0, n_in + 1, chunksize, dtype=np.int64)
    if isinstance(elev_azim_functions, tuple) and len(elev_azim_functions) == 3 and not callable(elev_azim_functions[0]):
        elev, azim, kinds = elev_azim_functions
    else:
        if len(elev_azim_functions) != n_src:
            raise ValueError('need one trajectory per source')
        per = [evaluate_trajectory(f, times) for f in elev_azim_functions]
        elev = np.stack([p[0] for p in per])
        azim = np.stack([p[1] for p in per])
        if all(np.isscalar(p[2]) for p in per) and len({p[2] for p in per}) == 1:
            kinds = per[0][2]
        else:
            kinds = np.stack([np.broadcast_to(np.asarray(p[2], dtype=np.uint8), (n_pts,)) for p in per])
    elev_d = torch.as_tensor(elev, dtype=torch.float64).to(device).contiguous()
    azim_d = torch.as_tensor(azim, dtype=torch.float64).to(device).contiguous()
    if elev_d.numel() != n_src * n_pts or azim_d.numel() != n_src * n_pts:
        raise ValueError('trajectories must give %d directions per source' % n_pts)

    filt, status, _ = _plan_and_synth(torch, dev, elev_d, azim_d, kinds, n_src * n_pts, _cabi.IR_ROWS)
    p0, p1 = (0, n_out) if time_range is None else (int(time_range[0]), int(time_range[1]))
    if not 0 <= p0 <= p1 <= n_out:
        raise ValueError('time_range outside [0, %d]' % n_out)
    count = p1 - p0
    stride = _round_up(max(count, 1), 4)
    out = torch.empty((1 if mix else n_src, 2, stride), dtype=torch.float32, device=device)
    peaks = torch.zeros(n_src, dtype=torch.float32, device=device)
    stream = _stream(torch)
    workspace = _cabi.render_workspace(torch, device)

    def launch(gains):
        _cabi.check(lib.bas_render(x.data_ptr(), n_in, n_in, n_src, n_in, chunksize, subchunksize, k,
                                   filt.data_ptr(), gains.data_ptr() if gains is not None else None,
                                   p0, count, out.data_ptr(), stride, 1 if mix else 0, peaks.data_ptr(), variant,
                                   workspace.data_ptr(), workspace.numel(), stream), 'bas_render')

    launch(None)
    if normalise and not mix:
        for s in range(n_src):        # apply_hrtf.py:462-464 per source, peak read on the device
            _cabi.check(lib.bas_normalise(out[s].data_ptr(), 2 * stride, peaks[s:s + 1].data_ptr(), stream), 'bas_normalise')
    # ---- results back: status + peaks and (unless return_device) the audio, one synchronisation ----
    small = torch.empty(2 + n_src, dtype=torch.int32, pin_memory=True)
    small.copy_(torch.cat([status, peaks.view(torch.int32)]), non_blocking=True)
    result = out[..., :count]
    if mix:
        result = result[0]
    host_out = None
    if not return_device:
        # pinned destination from torch's caching host allocator: DMA straight into the array the
        # caller receives (no pageable bounce, no extra host copy)
        host_out = torch.empty(result.shape, dtype=torch.float32, pin_memory=True)
        host_out.copy_(result, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    host = small.numpy()
    err, where = int(host[0]), int(host[1])
    if err:
        _raise_plan_error(err, ' (trajectory point %d of source %d)' % (where % n_pts, where // n_pts))
    peaks_host = host[2:].view(np.float32).copy()
    if normalise and mix and (peaks_host > 1).any():
        gains = torch.from_numpy((1.0 / np.maximum(peaks_host, 1.0)).astype(np.float32)).to(device)
        peaks.zero_()
        launch(gains)
        if host_out is not None:
            host_out.copy_(result, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    if host_out is not None:
        result = host_out.numpy()
    return (result, peaks_host) if return_peaks else result


def make_signal_move_2d(in_signal, chunksize: int, subchunksize: int, elev_azim_function, irs_and_delaydiffs):
    """apply_hrtf.py:356-466.  Filters the mono `in_signal` with HRIRs that follow
    elev_azim_function(t) (t in samples -> (elev, azim) radians): one interpolated HRIR pair per
    chunk boundary, linearly cross-faded per subchunk.  Returns float32 (N_out, 2), N_out =
    ceil(N / chunksize) * chunksize + K - 1, divided by its peak when that exceeds 1."""
    assert len(in_signal.shape) == 1, 'only mono signals for now'                   # apply_hrtf.py:398
    render_geometry(in_signal.shape[0], chunksize, subchunksize, irs_and_delaydiffs)   # :401-402 assertion first
    signals = in_signal[None, :]
    out = render_sources(signals, chunksize, subchunksize, [elev_azim_function], irs_and_delaydiffs)
    if PROGRESS:
        print(' 100.0%      ')                                                      # :457
    return out[0].T                                                                 # (N_out, 2), planar memory like :459
