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

import ctypes as C
import threading
import weakref

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
        # the same bank with every phase row stored twice in a row (+ padding): the layout bas_render_fused
        # gathers from ((m - adv) mod K becomes the plain index m + K - adv).  A copy, not arithmetic.
        self.bank_pp2 = torch.zeros(int(lib.bas_bank2_floats(u, self.taps)), dtype=torch.float32, device=device)
        self.bank_pp2[:2 * _cabi.N_DIRECTIONS * u * 2 * self.taps].view(2, _cabi.N_DIRECTIONS, u, 2, self.taps).copy_(
            self.bank_pp.unsqueeze(3).expand(2, _cabi.N_DIRECTIONS, u, 2, self.taps))
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
    mark = _bank_fingerprint(bank)
    state = cache.get(device.index)
    if state is None or state.fingerprint != mark:
        state = _BankDevice(bank, torch, device)          # first use, or the caller changed the bank's arrays
        state.fingerprint = mark
        cache[device.index] = state
    return state


def _bank_fingerprint(bank):
    """Cheap identity of a bank's contents: array identities, shapes and a strided sample of every
    array (a few hundred values), so that the cached device copy is dropped when the caller rebinds
    or overwrites the arrays.  Costs microseconds; not a cryptographic guarantee."""
    parts = [int(bank.upsampling)]
    for name in ('irs_left', 'irs_right', 'diffs_left', 'diffs_right'):
        a = np.asarray(getattr(bank, name))
        parts += [a.shape, a.__array_interface__['data'][0], float(a[::17, ::61].sum()) if a.ndim == 2 else float(a.sum())]
    return tuple(parts)


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
    """Raise what the reference raises for the error bits of ONE trajectory point (the earliest failing
    one, _cabi.decode_status), in the order the reference reaches its checks: the azimuth assertion of
    the ring lookups (sphere.py:87, called at apply_hrtf.py:214) comes before the vertical weight
    (:266) and before any int(floor(nan)) (:149).  Unlike the reference, which raises in the middle of
    its chunk loop, the exception surfaces after the device work of the call has been enqueued."""
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
    # (1 - alpha) is evaluated in alpha's own precision by the reference (apply_hrtf.py:90-91)
    terms_dev, delays_dev, status = _ring_plans(torch, dev, [(before, after, alpha)])
    width = dev.length if return_upsampled else dev.taps
    out = torch.empty((1, 2, width), dtype=torch.float32, device=dev.device)
    _cabi.check(lib.bas_ir_synth(dev.bank_pp.data_ptr(), dev.upsampling, dev.taps, terms_dev.data_ptr(), 1,
                                 _cabi.IR_UPSAMPLED if return_upsampled else _cabi.IR_PLANAR, out.data_ptr(), width,
                                 _stream(torch)), 'bas_ir_synth')
    _raise_plan_error(_cabi.decode_status(status.cpu())[0])
    delays = delays_dev.cpu().numpy()
    irs = out[0].cpu().numpy().astype(np.float64)
    return (delays[0, 0], delays[0, 1], irs)


def _ring_plans(torch, dev, triples):
    """bas_plan_ring for a list of (before, after, alpha): device tensors (terms, delays n x 2, status)."""
    n = len(triples)
    rows = np.empty((n, 2), dtype=np.int32)
    weights = np.empty((n, 2), dtype=np.float64)
    for i, (before, after, alpha) in enumerate(triples):
        rows[i] = (_grid_row(before), _grid_row(after))
        weights[i] = (float(alpha), float(1 - alpha))
    rows_d = torch.from_numpy(rows).to(dev.device)
    weights_d = torch.from_numpy(weights).to(dev.device)
    terms = torch.empty(n * 2 * _cabi.MAX_TERMS * 8, dtype=torch.uint8, device=dev.device)
    delays = torch.empty((n, 2), dtype=torch.float64, device=dev.device)
    status = torch.empty(2, dtype=torch.int32, device=dev.device)
    _cabi.check(lib.bas_plan_ring(dev.diffs[0].data_ptr(), dev.diffs[1].data_ptr(), dev.upsampling, dev.length, rows_d.data_ptr(),
                                  weights_d.data_ptr(), n, terms.data_ptr(), delays.data_ptr(), status.data_ptr(), _stream(torch)),
                'bas_plan_ring')
    return terms, delays, status


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


def delay_signal_float(in_sig, samples: float, downsample=1):
    """apply_hrtf.py:127-165: delay `in_sig` circularly by a non-integer number of samples (linear
    interpolation between the two neighbouring integer delays, np.roll semantics), then keep every
    `downsample`-th sample.  float64 in, float64 out, bit-identical to the reference (the blend runs
    in IEEE double on the device without fused multiply-adds).  Inside interpolate_2d the same
    operation is fused into bas_ir_synth; this entry point exists for callers that import it."""
    torch = _cabi.require_device()
    before = int(np.floor(samples))                                 # :149
    after = int(np.ceil(samples))                                   # :150
    a = samples - before                                            # :151
    x = np.ascontiguousarray(in_sig, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError('delay_signal_float takes a 1-D signal')
    n = x.size
    d = int(downsample) if downsample > 1 else 1                    # :160
    if n == 0:
        return np.zeros(0, dtype=np.float64)
    device = torch.device('cuda', torch.cuda.current_device())
    x_d = torch.from_numpy(x).to(device)
    out = torch.empty((n + d - 1) // d, dtype=torch.float64, device=device)
    _cabi.check(lib.bas_delay_signal_float(x_d.data_ptr(), n, before, after, float(a), d, out.data_ptr(), _stream(torch)),
                'bas_delay_signal_float')
    return out.cpu().numpy()


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
        err, where = _cabi.decode_status(status.cpu())
        if err:
            _raise_plan_error(err, ' (direction %d)' % where)
    if return_trace:
        return filt, trace.cpu().numpy().view(_cabi.TRACE_DTYPE).reshape(n)
    return filt


def _plan_and_synth(torch, dev, elev_d, azim_d, az_kind, n, mode, want_trace=False, status=None):
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
    if status is None:
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
    dirs = torch.tensor([float(elev), float(azim)], dtype=torch.float64).to(dev.device)
    filt, status, _ = _plan_and_synth(torch, dev, dirs[0:1], dirs[1:2], kind, 1, _cabi.IR_PLANAR)
    _raise_plan_error(_cabi.decode_status(status.cpu())[0])
    return filt[0].cpu().numpy().astype(np.float64)


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


# Trajectories.  The reference calls elev_azim_function(t) once per chunk boundary with a Python int
# (apply_hrtf.py:429, :435): 5,169 Python calls for a 60 s source, 24 ms - forty times the rest of the
# call.  Most trajectories are arithmetic on t that works unchanged on an array (every lambda in the
# reference's main does, apply_hrtf.py:583-593), so a plain callable is TRIAL-VECTORISED: one scalar
# call fixes the scalar type of the azimuth (it selects the arithmetic of the ring lookup, SURVEY.md
# section 5), one call with the whole int64 array of boundaries follows, and the array result is
# accepted only if it equals scalar calls at TRAJECTORY_CHECKS spot-check boundaries bit for bit
# (values and scalar kinds).  Anything else - an exception, a wrong shape, a mismatch - falls back to
# the reference's loop.  A callable can opt out (`fn.vectorized = False`) or declare itself array-safe
# and skip the checks (`fn.vectorized = True`, optionally `fn.az_kind`).
TRAJECTORY_TRIAL_MIN = 24        # fewer boundaries than this: just loop
TRAJECTORY_CHECKS = 8            # scalar spot checks of the first array evaluation of a callable (first, last, random)
TRAJECTORY_RECHECKS = 2          # spot checks of later evaluations (later phases of the same call): 12 per 3-phase call
TRAJECTORY_CHECKS_MANY = 4       # per callable when 16 or more sources are rendered in one call (first, last, two random)


def _scalar_point(elev_azim_function, t):
    e, a = elev_azim_function(int(t))                # a Python int, like range() gives the reference
    return float(e), float(a), sphere.az_kind(a)


def _try_vectorized(elev_azim_function, times, n_checks, rng):
    """(elev, azim, kind) from one array call, or None if the callable cannot be trusted with arrays."""
    n = len(times)
    try:
        e0, a0, kind = _scalar_point(elev_azim_function, times[0])
        res = elev_azim_function(np.asarray(times, dtype=np.int64))
        if not isinstance(res, tuple) or len(res) != 2:
            return None
        elev = np.asarray(res[0])
        azim = np.asarray(res[1])
        if elev.dtype.kind not in 'fiu' or azim.dtype.kind not in 'fiu' or elev.ndim > 1 or azim.ndim > 1:
            return None
        # an array azimuth behaves like its scalars: float32 arrays -> np.float32 scalars, everything else
        # -> the kind the scalar call showed (Python float and np.float64 compute the same float64 values)
        if (azim.dtype == np.float32) != (kind == _cabi.AZ_F32):
            return None
        elev = np.ascontiguousarray(np.broadcast_to(elev.astype(np.float64), (n,)))
        azim = np.ascontiguousarray(np.broadcast_to(azim.astype(np.float64), (n,)))
        picks = [0, n - 1] + [int(i) for i in rng.integers(0, n, max(0, n_checks - 2))]
        for i in picks:
            e, a, kd = (e0, a0, kind) if i == 0 else _scalar_point(elev_azim_function, times[i])
            same = (e == elev[i] or (e != e and elev[i] != elev[i])) and (a == azim[i] or (a != a and azim[i] != azim[i]))
            if not same or kd != kind:
                return None
        return elev, azim, int(kind)
    except Exception:
        return None


def evaluate_trajectory(elev_azim_function, times, state=None):
    """Directions at the chunk boundaries `times` (t = 0, C, ..., N_in; apply_hrtf.py:429, :435).
    Returns (elev float64[n], azim float64[n], kinds uint8[n] or a single kind).  `state` (a dict)
    carries the outcome of the trial vectorisation between the phases of one call."""
    n = len(times)
    declared = getattr(elev_azim_function, 'vectorized', None)
    if declared is True:
        elev, azim = elev_azim_function(np.asarray(times, dtype=np.int64))
        azim = np.asarray(azim)
        kind = getattr(elev_azim_function, 'az_kind', None)
        if kind is None:
            kind = _cabi.AZ_F32 if azim.dtype == np.float32 else _cabi.AZ_F64
        elev = np.broadcast_to(np.asarray(elev, dtype=np.float64), (n,))
        azim = np.broadcast_to(azim.astype(np.float64), (n,))
        return np.ascontiguousarray(elev), np.ascontiguousarray(azim), int(kind)
    state = {} if state is None else state
    if declared is None and n >= TRAJECTORY_TRIAL_MIN and state.get('mode') != 'loop':
        first = 'mode' not in state
        rng = state.setdefault('rng', np.random.default_rng(n))
        res = _try_vectorized(elev_azim_function, times, state.get('checks', TRAJECTORY_CHECKS) if first else TRAJECTORY_RECHECKS, rng)
        if res is not None and (first or res[2] == state.get('kind')):
            state['mode'], state['kind'] = 'array', res[2]
            return res
        state['mode'] = 'loop'
    elev = np.empty(n, dtype=np.float64)
    azim = np.empty(n, dtype=np.float64)
    kinds = np.empty(n, dtype=np.uint8)
    for i, t in enumerate(times):
        elev[i], azim[i], kinds[i] = _scalar_point(elev_azim_function, t)
    if n and (kinds == kinds[0]).all():
        return elev, azim, int(kinds[0])
    return elev, azim, kinds


# Host arrays in, host arrays out: the render is cut into time segments of about this many output
# bytes and upload / render / download of consecutive segments overlap on three streams.
PIPELINE_SEGMENT_BYTES = 8 << 20
PIPELINE_MAX_SEGMENTS = 64
_SEGMENT_ALIGN = 8192            # output samples; a whole number of render tiles for every tile width

_side_streams = {}
POISON_SCRATCH = False           # tests: fill freshly allocated device scratch with NaN, so that any read of an
                                 # uninitialised sample shows up in the output instead of depending on allocator history


def _scratch(torch, shape, dtype, device):
    t = torch.empty(shape, dtype=dtype, device=device)
    if POISON_SCRATCH:
        t.view(torch.uint8).fill_(0xff)
    return t



def _streams(torch, device):
    """(upload, download) side streams of `device`, created once."""
    st = _side_streams.get(device.index)
    if st is None:
        st = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
        _side_streams[device.index] = st
    return st


def _segments(p0: int, p1: int, bytes_per_sample: int):
    """Cut output samples [p0, p1) into pipeline segments at multiples of _SEGMENT_ALIGN."""
    count = p1 - p0
    if count <= 0:
        return []
    n_seg = max(1, min(PIPELINE_MAX_SEGMENTS, int(round(count * bytes_per_sample / PIPELINE_SEGMENT_BYTES))))
    step = _round_up(-(-count // n_seg), _SEGMENT_ALIGN)
    cuts = [p0]
    while cuts[-1] < p1:
        nxt = (cuts[-1] // _SEGMENT_ALIGN) * _SEGMENT_ALIGN + step
        cuts.append(min(p1, nxt))
    return list(zip(cuts[:-1], cuts[1:]))


def _host_directions(elev_azim_functions) -> bool:
    """True unless the caller passed pre-evaluated directions that already live on the device."""
    if isinstance(elev_azim_functions, tuple) and len(elev_azim_functions) == 3 and not callable(elev_azim_functions[0]):
        return not any(hasattr(v, 'is_cuda') and v.is_cuda for v in elev_azim_functions[:2])
    return True


def _directions(elev_azim_functions, n_src: int, n_in: int, chunksize: int):
    """(elev, azim, kinds) at t = 0, C, ..., N_in for every source: arrays of n_src x (N_in/C + 1)
    entries (numpy, or CUDA tensors when the caller passed those) and one az kind or an array."""
    n_pts = n_in // chunksize + 1
    if isinstance(elev_azim_functions, tuple) and len(elev_azim_functions) == 3 and not callable(elev_azim_functions[0]):
        return elev_azim_functions
    if len(elev_azim_functions) != n_src:
        raise ValueError('need one trajectory per source')
    times = np.arange(0, n_in + 1, chunksize, dtype=np.int64)
    per = [evaluate_trajectory(f, times) for f in elev_azim_functions]
    elev = per[0][0][None, :] if n_src == 1 else np.stack([p[0] for p in per])
    azim = per[0][1][None, :] if n_src == 1 else np.stack([p[1] for p in per])
    if all(np.isscalar(p[2]) for p in per) and len({p[2] for p in per}) == 1:
        kinds = per[0][2]
    else:
        kinds = np.stack([np.broadcast_to(np.asarray(p[2], dtype=np.uint8), (n_pts,)) for p in per])
    return elev, azim, kinds


# Pageable input arrays.  The reference's caller passes an ordinary ndarray (apply_hrtf.py:633); a
# DMA engine can only read page-locked memory, and a staged copy of a 60 s signal costs more than the
# rest of the call.  A large float32 array that owns its buffer is therefore page-locked IN PLACE on
# first sight (cudaHostRegister, about the cost of one staged copy) and stays registered for as long as
# the array object lives: a weakref finaliser releases it when the array dies, so repeated calls on the
# same signal upload by direct DMA.  Arrays that do not qualify take the driver's staged copy.
REGISTER_MIN_BYTES = 1 << 20
REGISTER_MAX_ARRAYS = 16
_registered = {}                 # id(owner) -> (address, bytes, finaliser), oldest first
_registered_lock = threading.Lock()


def _release_registration(key, address):
    with _registered_lock:
        _registered.pop(key, None)
    lib.bas_host_unregister(address)


def _pin_in_place(arr) -> bool:
    """Page-lock the buffer behind numpy array `arr` (see above); True if it is registered now."""
    owner = arr
    while isinstance(owner.base, np.ndarray):
        owner = owner.base
    if owner.base is not None or not owner.flags.owndata or owner.nbytes < REGISTER_MIN_BYTES:
        return False
    key, address, nbytes = id(owner), owner.ctypes.data, owner.nbytes
    with _registered_lock:
        hit = _registered.get(key)
        if hit is not None and hit[0] == address and hit[1] == nbytes:
            return True
        stale = [hit[2]] if hit is not None else []                  # the array was resized in place
        while len(_registered) - len(stale) >= REGISTER_MAX_ARRAYS:
            oldest = next(k for k in _registered if k != key)
            stale.append(_registered.pop(oldest)[2])
    for fin in stale:
        fin()                                                        # runs _release_registration now
    if lib.bas_host_register(address, nbytes) != 0:
        return False
    fin = weakref.finalize(owner, _release_registration, key, address)
    fin.atexit = False                                               # the CUDA runtime may be gone by then
    with _registered_lock:
        _registered[key] = (address, nbytes, fin)
    return True


class _HostCache(threading.local):
    """Per-thread, per-device scratch of the host pipeline, grown on demand and reused between calls
    (every call ends with its device work complete, so nothing is in flight when they are reused)."""

    def __init__(self):
        self.arena, self.pinned = {}, {}

    def device_arena(self, torch, device, nbytes):
        a = self.arena.get(device.index)
        if a is None or a.numel() < nbytes:
            a = None
            self.arena[device.index] = None
            a = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
            self.arena[device.index] = a
        if POISON_SCRATCH:
            a.fill_(0xff)
        return a

    def pinned_bytes(self, torch, name, nbytes):
        b = self.pinned.get(name)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes * 1.25), 4096), dtype=torch.uint8, pin_memory=True)
            self.pinned[name] = b
        return b


_host_cache = _HostCache()


# Phases of the host pipeline: cumulative fractions of the output range.  The trajectory of phase
# i+1 is evaluated on the host while phase i renders and travels back; a small first phase gets the
# device -> host copy started early.  Jobs below PIPELINE_PHASE_MIN_BYTES run as one phase.
PIPELINE_PHASES = (0.15, 0.5, 1.0)
# many sources: the signal upload (4 B per source sample) is the long pole, not the download (8 B per MIXED pair);
# more, shorter phases leave little to render after the last upload byte has arrived
PIPELINE_PHASES_MANY = (0.08, 0.2, 0.35, 0.5, 0.65, 0.8, 0.92, 1.0)
PIPELINE_PHASE_MIN_BYTES = 4 << 20


def _phase_plan(p0: int, p1: int, n_in: int, chunksize: int, out_bytes_per_sample: int, n_src: int = 1):
    """[(p_from, p_to, pt_begin, pt_end)] per phase: output range and the chunk-boundary directions
    that must be planned before it renders (every input below p_to needs both filters of its chunk)."""
    n_pts = n_in // chunksize + 1
    count = p1 - p0
    fracs = PIPELINE_PHASES if count * out_bytes_per_sample >= PIPELINE_PHASE_MIN_BYTES else (1.0,)
    if n_src >= 4 and count * 4 * n_src >= 4 * PIPELINE_PHASE_MIN_BYTES:
        fracs = PIPELINE_PHASES_MANY
    cuts = [p0]
    for f in fracs[:-1]:
        c = min(p1, _round_up(p0 + int(f * count), _SEGMENT_ALIGN))
        if c > cuts[-1]:
            cuts.append(c)
    if cuts[-1] < p1 or len(cuts) == 1:
        cuts.append(p1)
    phases, pt_done = [], 0
    for i in range(len(cuts) - 1):
        last = i == len(cuts) - 2
        last_in = min(cuts[i + 1], n_in) - 1
        pt_end = n_pts if last else min(n_pts, max(pt_done, last_in // chunksize + 2))
        phases.append((cuts[i], cuts[i + 1], pt_done, pt_end))
        pt_done = pt_end
    return phases


# Filter rows: synthesised inside the render kernel by its producer warps (bas_render_fused), or written to HBM
# first by a separate bas_ir_synth launch (the two-kernel path) - the same rows, bit for bit.  'auto' fuses wherever a
# fused tile shape fits shared memory (bas_render_fused_fits): measured on B200 the fused
# step is ahead at every tap count tried - per 60 s source, mixing, plan included: K = 100: 38 against 45 us,
# K = 256: 65 against 80, K = 384: 109 against 117, K = 512 (U = 16): 143 against 153.  True / False force either.
FUSED = 'auto'
FUSED_MAX_TAPS = 4096


def _want_fused(taps: int) -> bool:
    return bool(FUSED) if FUSED in (True, False) else taps <= FUSED_MAX_TAPS



class DeviceRender:
    """Sources resident in HBM, planned once, rendered range by range: the device half of
    render_sources, also driven by distributed.py (which interleaves collectives with the ranges).

    x: CUDA float32 (n_src, n_in) zero padded; elev_d / azim_d: CUDA float64 n_src * n_pts directions.
    plan() enqueues memset + bas_plan_build (+ bas_ir_synth unless fused) through bas_render_step;
    render(pa, pb, out_ptr, out_stride) enqueues the render of output samples [pa, pb)."""

    def __init__(self, torch, dev, x, n_in, chunksize, subchunksize, elev_d, azim_d, kinds, mix, variant):
        self.torch, self.dev, self.x = torch, dev, x
        device = dev.device
        n_src = x.shape[0]
        n_pts = n_in // chunksize + 1
        self.n_src, self.n_pts, self.mix = n_src, n_pts, mix
        self.elev_d, self.azim_d = elev_d, azim_d
        self.kinds_d = None if np.isscalar(kinds) else torch.as_tensor(
            np.ascontiguousarray(kinds, dtype=np.uint8).reshape(-1)).to(device)
        if self.kinds_d is not None and self.kinds_d.numel() != n_src * n_pts:
            raise ValueError('az_kind must have one entry per direction')
        self.fused = bool(_want_fused(dev.taps) and lib.bas_render_fused_fits(dev.taps, chunksize, subchunksize, 1 if mix else 0, variant)
                          and x.data_ptr() % 16 == 0 and (n_src == 1 or x.stride(0) % 4 == 0))
        self.terms = torch.empty(n_src * n_pts * 2 * _cabi.MAX_TERMS * 8, dtype=torch.uint8, device=device)
        self.filt = None if self.fused else torch.empty((n_src * n_pts, lib.bas_filter_row_pitch(dev.taps), 2),
                                                        dtype=torch.float32, device=device)
        self.small = torch.empty(2 + n_src, dtype=torch.int32, device=device)      # zeroed by plan()
        self.peaks = self.small[2:].view(torch.float32)
        self.workspace = _cabi.render_workspace(torch, device)
        self.job = _cabi.StepJob(
            n_src=n_src, C=chunksize, S=subchunksize, K=dev.taps, U=dev.upsampling, mix=1 if mix else 0, variant=variant,
            az_kind_all=_cabi.AZ_F64 if self.kinds_d is not None else int(kinds), flags=0,
            n_valid=n_in, n_in=n_in, x_stride=x.stride(0) if n_src > 1 else n_in, p_begin=0, p_count=0, out_stride=0,
            x_dev=x.data_ptr(), elev_dev=elev_d.data_ptr(), azim_dev=azim_d.data_ptr(),
            az_kind_dev=self.kinds_d.data_ptr() if self.kinds_d is not None else None,
            diffs_left_dev=dev.diffs[0].data_ptr(), diffs_right_dev=dev.diffs[1].data_ptr(), bank_pp_dev=dev.bank_pp.data_ptr(),
            bank_pp2_dev=dev.bank_pp2.data_ptr(), terms_dev=self.terms.data_ptr(),
            filt_dev=self.filt.data_ptr() if self.filt is not None else None, gains_dev=None, out_dev=None,
            small_dev=self.small.data_ptr(), workspace_dev=self.workspace.data_ptr(), workspace_bytes=self.workspace.numel())
        self._fused_flag = _cabi.STEP_FUSED if self.fused else 0

    def plan(self, stream):
        self.job.flags = _cabi.STEP_PLAN | self._fused_flag
        self.job.p_count = 0
        _cabi.check(lib.bas_render_step(C.byref(self.job), stream), 'bas_render_step')

    def render(self, stream, pa, pb, out_ptr, out_stride, gains=None, accumulate=False, normalise=False, route=None):
        """route (a _cabi.Route, mixing only): finished tiles go to their owner rank's receive buffer (distributed.PeerMix)."""
        job = self.job
        job.route = C.pointer(route) if route is not None else None
        job.flags = _cabi.STEP_RENDER | self._fused_flag | (_cabi.STEP_NORMALISE if normalise else 0)
        job.p_begin, job.p_count, job.out_dev, job.out_stride = pa, pb - pa, out_ptr, out_stride
        job.gains_dev = gains.data_ptr() if gains is not None else None
        job.mix = (_cabi.MIX_ACCUMULATE if accumulate else 1) if self.mix else 0
        rc = lib.bas_render_step(C.byref(job), stream)
        if rc == _cabi.E_UNSUPPORTED and self.fused:
            # no tile shape of the fused kernel fits this geometry: write the filter rows to HBM after all
            # and let bas_render choose (its generic kernel takes any chunk / subchunk / tap count)
            self._unfuse(stream)
            job.flags &= ~_cabi.STEP_FUSED
            rc = lib.bas_render_step(C.byref(job), stream)
        _cabi.check(rc, 'bas_render_step')

    def _unfuse(self, stream):
        dev, torch = self.dev, self.torch
        self.fused, self._fused_flag = False, 0
        self.filt = torch.empty((self.n_src * self.n_pts, lib.bas_filter_row_pitch(dev.taps), 2), dtype=torch.float32, device=dev.device)
        self.job.filt_dev = self.filt.data_ptr()
        _cabi.check(lib.bas_ir_synth(dev.bank_pp.data_ptr(), dev.upsampling, dev.taps, self.terms.data_ptr(), self.n_src * self.n_pts,
                                     _cabi.IR_ROWS, self.filt.data_ptr(), dev.taps, stream), 'bas_ir_synth')

    def zero_peaks(self):
        self.peaks.zero_()


def _render_pipeline(torch, dev, src, chunksize, subchunksize, elev_azim_functions, mix, normalise, variant,
                     p0, p1, return_peaks):
    """render_sources for host signals and a host result: lib.bas_pipeline_upload / _phase."""
    device = dev.device
    n_src, n = src.shape
    k = dev.taps
    n_in = _round_up(n, chunksize)
    n_pts = n_in // chunksize + 1
    n_dirs = n_src * n_pts
    count = p1 - p0
    n_rows = 1 if mix else n_src
    main = torch.cuda.current_stream()
    up, down = _streams(torch, device)
    offsets = (C.c_longlong * 8)()
    need = lib.bas_pipeline_arena_bytes(n_src, n_in, chunksize, k, 1 if mix else 0, count, 0, offsets)
    if need < 0:
        raise BasError('bas_pipeline_arena_bytes: bad geometry')
    arena = _host_cache.device_arena(torch, device, need)
    workspace = _cabi.render_workspace(torch, device)
    staged = _host_cache.pinned_bytes(torch, 'dirs', n_dirs * 17)
    small = _host_cache.pinned_bytes(torch, 'small', 4 * (2 + n_src))
    host_out = torch.empty((n_rows, 2, count), dtype=torch.float32, pin_memory=True)
    job = _cabi.PipelineJob(
        n_src=n_src, C=chunksize, S=subchunksize, K=k, U=dev.upsampling, mix=1 if mix else 0, variant=variant,
        az_kind_all=_cabi.AZ_F64, n=n, n_in=n_in, p_begin=p0, p_count=count, x_host_stride=n,
        segment_bytes=PIPELINE_SEGMENT_BYTES, x_host=src.data_ptr(), x_dev=None, dirs_host=staged.data_ptr(),
        az_kind_host=staged.data_ptr() + n_dirs * 16, diffs_left_dev=dev.diffs[0].data_ptr(),
        diffs_right_dev=dev.diffs[1].data_ptr(), bank_pp_dev=dev.bank_pp.data_ptr(), out_host=host_out.data_ptr(),
        small_host=small.data_ptr(), arena_dev=arena.data_ptr(), arena_bytes=arena.numel(),
        workspace_dev=workspace.data_ptr(), workspace_bytes=workspace.numel(),
        stream_main=main.cuda_stream, stream_up=up.cuda_stream, stream_down=down.cuda_stream,
        bank_pp2_dev=dev.bank_pp2.data_ptr() if _want_fused(k) else None)
    fused = bool(_want_fused(k) and lib.bas_render_fused_fits(k, chunksize, subchunksize, 1 if mix else 0, variant))
    phases = _phase_plan(p0, p1, n_in, chunksize, 8 * n_rows, n_src)
    cuts = (C.c_longlong * (len(phases) + 1))(*([ph[0] for ph in phases] + [p1]))
    _cabi.check(lib.bas_pipeline_upload(C.byref(job), len(phases), cuts), 'bas_pipeline_upload')
    staged_np = staged.numpy()
    elev_h = staged_np[:n_dirs * 8].view(np.float64).reshape(n_src, n_pts)
    azim_h = staged_np[n_dirs * 8:n_dirs * 16].view(np.float64).reshape(n_src, n_pts)
    kinds_h = staged_np[n_dirs * 16:n_dirs * 17].reshape(n_src, n_pts)
    pre = None
    if isinstance(elev_azim_functions, tuple) and len(elev_azim_functions) == 3 and not callable(elev_azim_functions[0]):
        pre = elev_azim_functions
        if np.size(pre[0]) != n_dirs or np.size(pre[1]) != n_dirs:
            raise ValueError('trajectories must give %d directions per source' % n_pts)
        elev_h[:] = np.asarray(pre[0], dtype=np.float64).reshape(n_src, n_pts)
        azim_h[:] = np.asarray(pre[1], dtype=np.float64).reshape(n_src, n_pts)
        kinds_h[:] = np.broadcast_to(np.asarray(pre[2], dtype=np.uint8), (n_src, n_pts)) if not np.isscalar(pre[2]) else int(pre[2])
    elif len(elev_azim_functions) != n_src:
        raise ValueError('need one trajectory per source')
    kinds_ptr = staged.data_ptr() + n_dirs * 16
    traj_state = [{} for _ in range(n_src)]
    if pre is None and n_src >= 4 and len(phases) > 2:
        # Many sources: a Python round trip per source and phase costs more than the phases' device work, and the
        # signal upload (the long pole) runs anyway.  Two evaluations per trajectory instead: the stretch the first
        # phase needs - so that rendering starts early - and all the rest while that phase uploads and renders;
        # one plan launch each.
        # Dozens of sources: the upload alone takes longer than everything the device has to do, so nothing is gained by
        # starting the first phase early - but the host must be through with all trajectories well before the last
        # sample has arrived.  ONE evaluation per trajectory (all points; one plan launch), with fewer spot checks.
        pt_a = phases[0][3] if n_src < 16 else n_pts
        phases = [(pa, pb, 0, pt_a) if i == 0 else (pa, pb, pt_a, n_pts) if i == 1 else (pa, pb, n_pts, n_pts)
                  for i, (pa, pb, _, _) in enumerate(phases)]
        if n_src >= 16:
            for st in traj_state:
                st['checks'] = TRAJECTORY_CHECKS_MANY          # dozens of callables: the spot checks are the host's long pole
    elif pre is not None:
        phases = [(pa, pb, 0 if i == 0 else n_pts, n_pts) for i, (pa, pb, _, _) in enumerate(phases)]
    for i, (pa, pb, pt0, pt1) in enumerate(phases):
        # the trajectory of this phase is evaluated while the signal and the earlier phases travel
        if pre is None and pt1 > pt0:
            times = np.arange(pt0 * chunksize, (pt1 - 1) * chunksize + 1, chunksize, dtype=np.int64)
            for s, fn in enumerate(elev_azim_functions):
                e, a, kd = evaluate_trajectory(fn, times, traj_state[s])
                elev_h[s, pt0:pt1] = e
                azim_h[s, pt0:pt1] = a
                kinds_h[s, pt0:pt1] = kd
        # one az kind for the whole phase (the usual case): the directions ride in the plan launches
        first_kind = int(kinds_h[0, pt0]) if pt1 > pt0 else _cabi.AZ_F64
        uniform = pt1 == pt0 or bool((kinds_h[:, pt0:pt1] == first_kind).all())
        job.az_kind_all = first_kind
        job.az_kind_host = None if uniform else kinds_ptr
        if PROGRESS and len(phases) > 1 and i:
            print(' {:.1f}%           '.format(100 * (pa - p0) / max(count, 1)), end='\r')      # apply_hrtf.py:456, once per phase
        _cabi.check(lib.bas_pipeline_phase(C.byref(job), i, len(phases), pt0, pt1, pa, pb), 'bas_pipeline_phase')
    words = small.numpy()[:4 * (2 + n_src)].view(np.int32)
    err, where = _cabi.decode_status(words)
    if err:
        _raise_plan_error(err, ' (trajectory point %d of source %d)' % (where % n_pts, where // n_pts))
    peaks_host = words[2:].view(np.float32).copy()
    if normalise and count > 0 and (peaks_host > 1).any():
        # apply_hrtf.py:462-464, the rare second pass: divide on the device, copy again
        stride = int(offsets[7])
        base = arena.data_ptr()
        out_dev, peaks_dev, stream = base + offsets[6], base + offsets[4] + 8, main.cuda_stream
        if mix:
            gains = torch.from_numpy((1.0 / np.maximum(peaks_host, 1.0)).astype(np.float32)).to(device)
            if fused:
                _cabi.check(lib.bas_render_fused(base + offsets[5], n_in, n_in, n_src, n_in, chunksize, subchunksize, k, base + offsets[2],
                                                 dev.bank_pp2.data_ptr(), dev.upsampling, gains.data_ptr(), p0, count, out_dev, stride, 1,
                                                 None, variant, workspace.data_ptr(), workspace.numel(), stream), 'bas_render_fused')
            else:
                _cabi.check(lib.bas_render(base + offsets[5], n_in, n_in, n_src, n_in, chunksize, subchunksize, k, base + offsets[3],
                                           gains.data_ptr(), p0, count, out_dev, stride, 1, None, variant, workspace.data_ptr(),
                                           workspace.numel(), stream), 'bas_render')
        else:
            for s in np.nonzero(peaks_host > 1)[0]:
                _cabi.check(lib.bas_normalise(out_dev + int(s) * 8 * stride, 2 * stride, peaks_dev + 4 * int(s), stream), 'bas_normalise')
        _cabi.check(lib.bas_copy_2d(host_out.data_ptr(), 4 * count, out_dev, 4 * stride, 4 * count, 2 * n_rows, 0, stream), 'bas_copy_2d')
        main.synchronize()
    result = host_out.numpy()
    if mix:
        result = result[0]
    return (result, peaks_host) if return_peaks else result


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
    return_device: return CUDA tensors instead of numpy arrays (no host copy).

    Host signals with a host result run as a pipeline over time segments: segment i+1 is uploaded
    and segment i-1 downloaded while segment i renders (PCIe is full duplex), so the call costs
    about one transfer of the larger direction instead of upload + render + download."""
    torch = _cabi.require_device()
    dev = _device_bank(irs_and_delaydiffs)
    device = dev.device
    main = torch.cuda.current_stream()
    stream = main.cuda_stream
    if isinstance(signals, torch.Tensor):
        # CUDA tensors are used where they lie; host tensors are converted like host arrays
        src = signals if signals.is_cuda else signals.detach().to(torch.float32).contiguous()
    else:
        host = np.ascontiguousarray(signals, dtype=np.float32)
        if not return_device:
            _pin_in_place(host)
        src = torch.from_numpy(host)
    if src.dim() != 2:
        raise ValueError('signals must be (n_src, N)')
    n_src, n = src.shape
    k, n_in, n_out = render_geometry(n, chunksize, subchunksize, irs_and_delaydiffs)
    assert k == dev.taps
    p0, p1 = (0, n_out) if time_range is None else (int(time_range[0]), int(time_range[1]))
    if not 0 <= p0 <= p1 <= n_out:
        raise ValueError('time_range outside [0, %d]' % n_out)
    count = p1 - p0
    n_pts = n_in // chunksize + 1
    if n == 0:
        # An empty signal: the reference still evaluates the trajectory at t = 0 and interpolates that
        # filter (apply_hrtf.py:429, with its assertions), runs no chunk, and returns K - 1 zero pairs.
        elev, azim, kinds = _directions(elev_azim_functions, n_src, 0, chunksize)
        interpolate_2d_batch(irs_and_delaydiffs, elev, azim, kinds)
        zeros = torch.zeros((2, count) if mix else (n_src, 2, count), dtype=torch.float32, device=device if return_device else 'cpu')
        result = zeros if return_device else zeros.numpy()
        return (result, np.zeros(n_src, dtype=np.float32)) if return_peaks else result
    if not src.is_cuda and not return_device and src.dtype == torch.float32 and _host_directions(elev_azim_functions):
        return _render_pipeline(torch, dev, src, chunksize, subchunksize, elev_azim_functions, mix, normalise, variant,
                                p0, p1, return_peaks)
    up, down = _streams(torch, device)

    # ---- signals: HBM-resident as they are, or uploaded (zero padded to a multiple of chunksize,
    #      apply_hrtf.py:405-406).  The upload is issued first, on its own stream: it overlaps the
    #      trajectory evaluation on the host and the plan / ir_synth kernels. --------------------------
    resident = src.is_cuda and src.dtype == torch.float32 and src.is_contiguous() and n == n_in and src.device == device
    uploaded = None
    if resident:
        x = src
    else:
        x = _scratch(torch, (n_src, n_in), torch.float32, device)
        if n_in > n:
            x[:, n:].zero_()
        if src.is_cuda:
            x[:, :n].copy_(src)
        elif count > 0:
            # first input sample the kernels READ (the tiled kernel multiplies samples back to
            # p0/32*32 - 32*ceil(K/32) by zero-padding taps: they must be initialised, 0 * NaN = NaN) and
            # one past the last sample any output needs
            lo = max(0, p0 // 32 * 32 - 32 * ((k + 31) // 32))
            hi = min(n, p1)
            up.wait_stream(main)                            # x was just allocated on `main`
            _cabi.check(lib.bas_copy_2d(x.data_ptr() + 4 * lo, 4 * n_in, src.data_ptr() + 4 * lo, 4 * n, 4 * (hi - lo), n_src, 1,
                                        up.cuda_stream), 'bas_copy_2d')
            uploaded = torch.cuda.Event()
            uploaded.record(up)

    # ---- directions at the chunk boundaries -> HBM (one pinned staging buffer, one copy) ----------
    elev, azim, kinds = _directions(elev_azim_functions, n_src, n_in, chunksize)
    if isinstance(elev, torch.Tensor) or isinstance(azim, torch.Tensor):
        elev_d = torch.as_tensor(elev, dtype=torch.float64).to(device).contiguous()
        azim_d = torch.as_tensor(azim, dtype=torch.float64).to(device).contiguous()
    else:
        if np.size(elev) != n_src * n_pts or np.size(azim) != n_src * n_pts:
            raise ValueError('trajectories must give %d directions per source' % n_pts)
        staged = torch.empty((2, n_src * n_pts), dtype=torch.float64, pin_memory=True)
        host = staged.numpy()
        host[0] = np.asarray(elev, dtype=np.float64).reshape(-1)
        host[1] = np.asarray(azim, dtype=np.float64).reshape(-1)
        both = staged.to(device, non_blocking=True)
        elev_d, azim_d = both[0], both[1]
    if elev_d.numel() != n_src * n_pts or azim_d.numel() != n_src * n_pts:
        raise ValueError('trajectories must give %d directions per source' % n_pts)
    stride = _round_up(max(count, 1), 4)
    n_rows = 1 if mix else n_src
    out = torch.empty((n_rows, 2, stride), dtype=torch.float32, device=device)
    job = DeviceRender(torch, dev, x, n_in, chunksize, subchunksize, elev_d.reshape(-1), azim_d.reshape(-1), kinds, mix, variant)
    job.plan(stream)
    if uploaded is not None:
        main.wait_event(uploaded)

    def launch(gains, pa, pb, normalise_now=False):
        job.render(stream, pa, pb, out.data_ptr() + 4 * (pa - p0), stride, gains=gains, normalise=normalise_now)

    small = torch.empty(2 + n_src, dtype=torch.int32, pin_memory=True)
    host_out = None if return_device else torch.empty((n_rows, 2, count), dtype=torch.float32, pin_memory=True)

    def download(pa, pb, st):
        """out[:, :, pa-p0 : pb-p0] -> host_out, 2 * n_rows rows of (pb - pa) floats."""
        _cabi.check(lib.bas_copy_2d(host_out.data_ptr() + 4 * (pa - p0), 4 * count, out.data_ptr() + 4 * (pa - p0), 4 * stride,
                                    4 * (pb - pa), 2 * n_rows, 0, st.cuda_stream), 'bas_copy_2d')

    def fetch_small(st):
        _cabi.check(lib.bas_copy_2d(small.data_ptr(), 0, job.small.data_ptr(), 0, 4 * (2 + n_src), 1, 0, st.cuda_stream), 'bas_copy_2d')

    if host_out is not None and count > 0:
        # ---- pipeline: segment i is downloaded while segment i+1 renders ------------------------------
        down.wait_stream(main)                              # out was just allocated on `main`
        for i, (pa, pb) in enumerate(_segments(p0, p1, 8 * n_rows)):
            launch(None, pa, pb)
            ev = torch.cuda.Event()
            ev.record(main)
            down.wait_event(ev)
            download(pa, pb, down)
        fetch_small(down)
        down.synchronize()
        second_pass = normalise
    else:
        if count > 0:
            launch(None, p0, p1, normalise_now=normalise and not mix)   # apply_hrtf.py:462-464 per source, peak read on the device
        second_pass = normalise and mix
        fetch_small(main)
        main.synchronize()

    host = small.numpy()
    err, where = _cabi.decode_status(host)
    if err:
        _raise_plan_error(err, ' (trajectory point %d of source %d)' % (where % n_pts, where // n_pts))
    peaks_host = host[2:].view(np.float32).copy()
    # ---- apply_hrtf.py:462-464: divide by the peak where it exceeds 1.  A second pass, needed only
    #      when a peak does exceed 1 and the audio has already left (pipeline) or been mixed --------
    if second_pass and count > 0 and (peaks_host > 1).any():
        if mix:
            gains = torch.from_numpy((1.0 / np.maximum(peaks_host, 1.0)).astype(np.float32)).to(device)
            job.zero_peaks()
            launch(gains, p0, p1)
        else:
            for s in np.nonzero(peaks_host > 1)[0]:
                _cabi.check(lib.bas_normalise(out[int(s)].data_ptr(), 2 * stride, job.peaks[int(s):int(s) + 1].data_ptr(), stream),
                            'bas_normalise')
        if host_out is not None:
            download(p0, p1, main)
        main.synchronize()
    result = host_out.numpy() if host_out is not None else out[..., :count]
    if mix:
        result = result[0]
    return (result, peaks_host) if return_peaks else result


def suggest_chunk_sizes(elev_azim_function, n_samples: int, max_chunk_degrees: float = 6.0, max_subchunk_degrees: float = 0.5,
                        chunk_sizes=(1024, 512, 256, 128), subchunk_sizes=(64, 32, 16)):
    """The reference's "idea for the future" (apply_hrtf.py:383-385): choose chunksize and subchunksize from how fast
    the source moves.  The trajectory is sampled every 128 samples; the pair returned is the largest chunksize whose
    fastest chunk turns the source by at most max_chunk_degrees (one HRIR interpolation per chunk: less than half a
    15-degree grid cell by default) and the largest subchunksize whose fastest subchunk turns it by at most
    max_subchunk_degrees (one cross-fade step per subchunk).  Every pair this returns runs on the tiled kernel
    (subchunksize 16, 32 or 64; chunksize a multiple of both).  One choice per call: sizes that vary inside a
    call would no longer be the reference's algorithm."""
    probe = 128
    times = np.arange(0, max(int(n_samples), probe) + 1, probe, dtype=np.int64)
    elev, azim, _ = evaluate_trajectory(elev_azim_function, times)
    elev = np.clip(elev, -np.pi / 4, np.pi / 2)              # the renderer clamps to the grid's elevation range
    u = np.stack([np.cos(elev) * np.cos(azim), np.cos(elev) * np.sin(azim), np.sin(elev)], axis=1)
    if len(u) < 2:
        return int(chunk_sizes[0]), int(subchunk_sizes[0])
    step = np.degrees(np.arccos(np.clip((u[1:] * u[:-1]).sum(axis=1), -1.0, 1.0)))       # degrees per 128 samples
    speed = float(np.nanmax(step)) / probe if np.isfinite(step).any() else 0.0               # degrees per sample, fastest stretch
    chunk = next((c for c in chunk_sizes if c * speed <= max_chunk_degrees), chunk_sizes[-1])
    sub = next((b for b in subchunk_sizes if b * speed <= max_subchunk_degrees and chunk % b == 0), subchunk_sizes[-1])
    return int(chunk), int(sub)


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


def make_signal_move(in_signal, chunksize: int, index_function, irs_and_delaydiffs):
    """apply_hrtf.py:294-353, the legacy 1-D renderer (superseded by make_signal_move_2d per its own
    docstring): index_function(t) gives a continuous index on the horizontal ring; every chunk is
    convolved with ONE ring-interpolated filter (delay_compensated_interpolation_easy, no cross-fade)
    and overlap-added.  Same kernels as the 2-D path: the scalar part of the ring interpolation of all
    chunks in one bas_plan_ring launch, bas_ir_synth for the filters, and bas_render with
    subchunksize = chunksize, for which every blend weight alpha_q is zero."""
    torch = _cabi.require_device()
    assert len(in_signal.shape) == 1, 'only mono signals for now'                   # apply_hrtf.py:308
    dev = _device_bank(irs_and_delaydiffs)
    device = dev.device
    k = int(0.5 + irs_and_delaydiffs.irs_left.shape[1] / irs_and_delaydiffs.upsampling)
    n = in_signal.size
    n_in = int(0.5 + np.ceil(n / chunksize) * chunksize)
    n_out = n_in + k - 1
    n_chunks = n_in // chunksize
    triples = []
    for c in range(n_chunks):
        ci = index_function(c * chunksize)                                          # :334
        before, after = int(np.floor(ci)), int(np.ceil(ci))                         # :118-119
        alpha = ci - before
        if after == 97:
            after = 73                                                              # :121-122
        triples.append((before, after, alpha))
    triples.append(triples[-1])                   # the boundary after the last chunk: present in the layout, weight 0
    terms_dev, _, status = _ring_plans(torch, dev, triples)
    stream = _stream(torch)
    filt = torch.empty((n_chunks + 1, lib.bas_filter_row_pitch(k), 2), dtype=torch.float32, device=device)
    _cabi.check(lib.bas_ir_synth(dev.bank_pp.data_ptr(), dev.upsampling, k, terms_dev.data_ptr(), n_chunks + 1, _cabi.IR_ROWS,
                                 filt.data_ptr(), k, stream), 'bas_ir_synth')
    x = torch.zeros(n_in, dtype=torch.float32, device=device)
    x[:n] = torch.from_numpy(np.ascontiguousarray(in_signal, dtype=np.float32)).to(device)
    stride = _round_up(n_out, 4)
    out = torch.empty((2, stride), dtype=torch.float32, device=device)
    peak = torch.zeros(1, dtype=torch.float32, device=device)
    variant = _cabi.RENDER_AUTO if chunksize == 32 else _cabi.RENDER_GENERIC
    _cabi.check(lib.bas_render(x.data_ptr(), n_in, n_in, 1, n_in, chunksize, chunksize, k, filt.data_ptr(), None, 0, n_out,
                               out.data_ptr(), stride, 0, peak.data_ptr(), variant, None, 0, stream), 'bas_render')
    _cabi.check(lib.bas_normalise(out.data_ptr(), 2 * stride, peak.data_ptr(), stream), 'bas_normalise')   # :349-351
    err, where = _cabi.decode_status(status.cpu())
    if err:
        _raise_plan_error(err, ' (chunk %d)' % where)
    if PROGRESS:
        print(' 100.0%      ')                                                      # :346
    return out[:, :n_out].cpu().numpy().T
