"""CPU oracle for the bank builder (upsample_irs.m).  TEST INFRASTRUCTURE ONLY.

float64 numpy restatements of the two array computations of `upsample_irs.m` that
`binaural-audio-synthesis_b200/csrc/bank_builder.cu` runs on the device: `resample(row, U, 1)`
(upsample_irs.m:42-43) for a given zero-phase FIR, and `delaydifference` (:59-77) with
`parabolic_interpolation` (:88-101).  Octave is not installed here, so these follow the published
semantics of the functions the .m file calls and are checked against scipy where scipy implements
the same thing (`resample_poly`); parity with Octave itself is unpinned.  Never imported by the
product package.
"""
import numpy as np


def resample(x: np.ndarray, p: int, h: np.ndarray) -> np.ndarray:
    """Zero-phase interpolation by p with the odd-length FIR h: y[m] = sum_k h[Lh + m - k p] x[k]."""
    x = np.asarray(x, dtype=np.float64)
    half = (h.size - 1) // 2
    stuffed = np.zeros(x.size * p)
    stuffed[::p] = x
    return np.convolve(stuffed, h)[half:half + x.size * p]


def delay_difference(a: np.ndarray, b: np.ndarray, p: int, h: np.ndarray) -> float:
    """delaydifference, upsample_irs.m:59-77 with parabolic_interpolation :88-101."""
    n = a.size
    cc = np.convolve(np.asarray(a, dtype=np.float64)[::-1], np.asarray(b, dtype=np.float64))      # fftconv(fliplr(a), b)
    up = resample(cc, p, h)
    pk = int(np.argmax(up))                                       # first maximum, 0-based
    v0, v1, v2 = up[pk - 1], up[pk], up[pk + 1]
    frac = -(0.5 * (v2 - v0)) / (2 * (0.5 * (v0 + v2 - 2 * v1)))
    return (pk + 1 + frac - 1) / p - (n - 1)
