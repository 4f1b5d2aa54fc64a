"""Synthetic HRIR bank in the on-disk layout `upsample_irs.m` produces.

The IRCAM LISTEN database the reference was built on (upsample_irs.m:1) is not available
offline, so benchmarks and tests use a synthetic bank on the same 187-direction grid
(sphere.py:127-315; elevation >= -45 deg).  This module restates the *offline* preprocessing
of `upsample_irs.m:15-54` in numpy so that the file it writes has exactly the struct the
loader expects (apply_hrtf.py:34-44):

    irs_and_delaydiffs.upsampling             double scalar          (upsample_irs.m:46-47)
    irs_and_delaydiffs.diffs_left/right       187 x 187 double       (:17-32, antisymmetric)
    irs_and_delaydiffs.irs_left/right         187 x (512*U) double   (:37-44)

It only defines input data; it is not on the parity path.
"""
from __future__ import annotations

import numpy as np

from . import _grid

BASE_LENGTH = 512   # samples per measured HRIR (upsample_irs.m:37-38)


def synthetic_hrirs(seed: int = 0, length: int = BASE_LENGTH):
    """(left, right) arrays of shape (187, length): a band-limited pulse at a direction-
    dependent fractional onset (spherical-head-like interaural delay) plus a decaying
    low-passed noise tail.  Neighbouring grid cells differ by at most a few samples of
    delay, like real HRIRs, so the delay-compensated interpolation is exercised properly."""
    rng = np.random.default_rng(seed)
    tab = _grid.get_index_elev_azim().astype(np.float64)
    elev, azim = tab[:, 1], tab[:, 2]
    n = np.arange(length, dtype=np.float64)
    lateral = np.sin(azim) * np.cos(elev)          # +1: source fully to the left
    ears = []
    for sign in (-1.0, +1.0):                      # left ear hears left sources earlier
        onset = 30.0 + sign * 14.0 * lateral
        gain = 1.0 - 0.35 * sign * lateral         # head shadow
        t = n[None, :] - onset[:, None]
        win = np.where(np.abs(t) < 12, 0.5 * (1 + np.cos(np.pi * t / 12)), 0.0)
        pulse = gain[:, None] * np.sinc(0.8 * t) * win
        noise = rng.standard_normal((tab.shape[0], length))
        # cheap one-pole low-pass keeps the tail band-limited well below Nyquist
        for k in range(1, length):
            noise[:, k] = 0.55 * noise[:, k - 1] + 0.45 * noise[:, k]
        tail = 0.12 * noise * np.exp(-np.clip(t, 0, None) / 40.0) * (t > 2)
        ears.append(pulse + tail)
    return ears[0], ears[1]


def _parabolic_peak(v_prev, v_mid, v_next):
    """Vertex of the parabola through three equally spaced points (upsample_irs.m:88-101)."""
    a = 0.5 * (v_prev + v_next - 2 * v_mid)
    b = 0.5 * (v_next - v_prev)
    return -b / (2 * a)


def delay_differences(irs: np.ndarray, upsampling: int, batch: int = 4096) -> np.ndarray:
    """187 x 187 antisymmetric matrix of pairwise delay differences in base-rate samples.

    upsample_irs.m:59-77 cross-correlates each pair, upsamples the correlation U times,
    takes the arg-max, refines it with a parabola and re-centres on zero lag; :22-32 fills
    the upper triangle and antisymmetrises.  Here the correlation is formed by FFT for all
    pairs and only a window around the coarse peak is upsampled (the polyphase filter is
    local), which gives the same peak for pulse-like responses.
    """
    from scipy.signal import resample_poly
    n_dir, length = irs.shape
    nfft = 2 * length
    spec = np.fft.rfft(irs, nfft, axis=1)
    ii, jj = np.triu_indices(n_dir, k=1)
    out = np.zeros((n_dir, n_dir))
    half = 32
    for s in range(0, ii.size, batch):
        i, j = ii[s:s + batch], jj[s:s + batch]
        cc = np.fft.irfft(np.conj(spec[i]) * spec[j], nfft, axis=1)
        cc = np.roll(cc, length - 1, axis=1)[:, :2 * length - 1]   # index = lag + (length-1)
        coarse = np.argmax(cc, axis=1)
        coarse = np.clip(coarse, half, cc.shape[1] - half - 1)
        cols = coarse[:, None] + np.arange(-half, half + 1)[None, :]
        win = np.take_along_axis(cc, cols, axis=1)
        up = resample_poly(win, upsampling, 1, axis=1)
        lo, hi = (half - 8) * upsampling, (half + 8) * upsampling
        peak = lo + np.argmax(up[:, lo:hi + 1], axis=1)
        r = np.arange(peak.size)
        frac = _parabolic_peak(up[r, peak - 1], up[r, peak], up[r, peak + 1])
        pos = (peak + frac) / upsampling + (coarse - half)       # position in the full correlation
        out[i, j] = pos - (length - 1)
    return out - out.T


def build_bank(upsampling: int = 8, seed: int = 0):
    """dict with the five fields of the reference's struct (upsample_irs.m:46-50)."""
    from scipy.signal import resample_poly
    left, right = synthetic_hrirs(seed)
    return {
        'upsampling': float(upsampling),
        'diffs_left': delay_differences(left, upsampling),
        'diffs_right': delay_differences(right, upsampling),
        'irs_left': resample_poly(left, upsampling, 1, axis=1),      # upsample_irs.m:42
        'irs_right': resample_poly(right, upsampling, 1, axis=1),    # upsample_irs.m:43
    }


def write_mat(filename: str, fields: dict) -> None:
    """Write the struct as a MATLAB v5 file without compression - what Octave's `save -6`
    emits (upsample_irs.m:53) and what apply_hrtf.py:34 reads."""
    import scipy.io
    scipy.io.savemat(filename, {'irs_and_delaydiffs': fields}, format='5', do_compression=False)


def write_synthetic_bank(filename: str, upsampling: int = 8, seed: int = 0) -> dict:
    fields = build_bank(upsampling, seed)
    write_mat(filename, fields)
    return fields
