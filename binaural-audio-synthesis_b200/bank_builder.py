"""Bank builder on the GPU: `upsample_irs.m` (the reference's offline preprocessing).

    upsample_irs(l_hrir, r_hrir, upsampling)  ->  the struct fields of upsample_irs.m:46-50
        irs_left/right   187 x (512 U)   resample(row, U, 1)                       (:42-43)
        diffs_left/right 187 x 187       delaydifference of every pair, diffs - diffs'   (:17-32, :59-77)

The two array computations run in libbas_b200.so (csrc/bank_builder.cu, fp64); the resampling
filter is designed here on the host and handed to the kernels, so either of two designs can be used:

    'octave'  the Kaiser-windowed sinc that Octave-signal's resample.m designs when called as
              resample(x, p, 1) - restated from its published algorithm (60 dB rejection, roll-off
              1/(20 p), length from Oppenheim & Schafer eq. 7.63).  Octave is not installed here and
              the signal package's version is not pinned by the reference, so this restatement is
              UNPINNED: it is checked against its own numpy twin only.
    'scipy'   scipy.signal.resample_poly's default (Kaiser beta = 5, half length 10 p), which is what
              bank_synth.py has used for the synthetic banks of the tests and benchmarks.

There is no CPU path here: the float64 numpy restatements the tests compare the kernels with live in
`oracle/bank_oracle.py`.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from ._cabi import lib


def octave_filter(p: int) -> np.ndarray:
    """FIR that Octave-signal's resample(x, p, 1) designs (see the module docstring)."""
    log10_rejection = -3.0
    stopband_cutoff_f = 1.0 / (2.0 * p)
    roll_off_width = stopband_cutoff_f / 10.0
    rejection_db = -20.0 * log10_rejection
    half = int(np.ceil((rejection_db - 8.0) / (28.714 * roll_off_width)))
    t = np.arange(-half, half + 1, dtype=np.float64)
    ideal = 2 * p * stopband_cutoff_f * np.sinc(2 * stopband_cutoff_f * t)
    if 21 <= rejection_db <= 50:
        beta = 0.5842 * (rejection_db - 21.0) ** 0.4 + 0.07886 * (rejection_db - 21.0)
    elif rejection_db > 50:
        beta = 0.1102 * (rejection_db - 8.7)
    else:
        beta = 0.0
    return np.kaiser(2 * half + 1, beta) * ideal


def scipy_filter(p: int) -> np.ndarray:
    """scipy.signal.resample_poly(x, p, 1)'s default filter."""
    from scipy.signal import firwin
    half = 10 * p
    return firwin(2 * half + 1, 1.0 / p, window=('kaiser', 5.0)) * p


def design_filter(p: int, kind: str = 'octave') -> np.ndarray:
    if kind == 'octave':
        return octave_filter(p)
    if kind == 'scipy':
        return scipy_filter(p)
    raise ValueError("filter must be 'octave' or 'scipy'")


def upsample_irs(l_hrir, r_hrir, upsampling: int, filter: str = 'octave', filename: str | None = None) -> dict:
    """upsample_irs.m:15-54 on the device.  l_hrir / r_hrir: (n_rows, n) measured HRIRs
    (`l_eq_hrir_S.content_m`, `r_eq_hrir_S.content_m`).  Returns the struct fields; with `filename`
    also writes them like `save -6` (:53)."""
    torch = _cabi.require_device()
    device = torch.device('cuda', torch.cuda.current_device())
    stream = torch.cuda.current_stream().cuda_stream
    u = int(upsampling)
    h = np.ascontiguousarray(design_filter(u, filter), dtype=np.float64)
    h_dev = torch.from_numpy(h).to(device)
    fields = {'upsampling': float(u)}
    for name, hrir in (('left', l_hrir), ('right', r_hrir)):
        x = np.ascontiguousarray(hrir, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError('HRIRs must be (n_rows, n)')
        n_rows, n = x.shape
        x_dev = torch.from_numpy(x).to(device)
        irs = torch.empty((n_rows, n * u), dtype=torch.float64, device=device)
        diffs = torch.empty((n_rows, n_rows), dtype=torch.float64, device=device)
        _cabi.check(lib.bas_bank_upsample(x_dev.data_ptr(), n_rows, n, u, h_dev.data_ptr(), h.size, irs.data_ptr(), stream), 'bas_bank_upsample')
        _cabi.check(lib.bas_bank_delay_diffs(x_dev.data_ptr(), n_rows, n, u, h_dev.data_ptr(), h.size, diffs.data_ptr(), stream),
                    'bas_bank_delay_diffs')
        fields['irs_' + name] = irs.cpu().numpy()
        fields['diffs_' + name] = diffs.cpu().numpy()
    if filename is not None:
        from .bank_synth import write_mat
        write_mat(filename, fields)
    return fields
