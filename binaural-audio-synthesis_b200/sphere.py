"""Measurement grid of the HRIR bank: drop-in for the hot-path part of the reference's sphere.py.

`index_elev_azim` reproduces the 187 x 3 float32 table (index, elevation, azimuth in radians) that
sphere.py:124-319 lists literally and converts to radians in float32 (sphere.py:318); here it is
generated from the ring structure (ten elevation rings, 24/12/6/1 points).  The lookup
`azim_to_interpolation_params` (sphere.py:78-121) runs in the C-ABI's host scalar helper
(plan_math.h), the same source the device plan kernel is compiled from.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi

from ._grid import RING_ELEV_DEG, RING_COUNT, get_index_elev_azim          # noqa: F401  (pure numpy, no library)

index_elev_azim = get_index_elev_azim()            # module global, like sphere.py:350


def az_kind(azim) -> int:
    """Which arithmetic the reference's sphere.py:86,103-119 would use for this azimuth object
    under NumPy-2 promotion: a Python scalar is 'weak' and is compared against
    the float32 table in float32; a float64 keeps float64; a float32 also takes the modulo in
    float32."""
    if isinstance(azim, np.ndarray):
        dt = azim.dtype
    elif isinstance(azim, np.generic):
        dt = azim.dtype
    else:
        return _cabi.AZ_PYFLOAT                    # Python float / int / bool
    if dt == np.float32 or dt == np.float16:
        return _cabi.AZ_F32
    return _cabi.AZ_F64                             # float64, longdouble, and integers (int % float -> float64)


def azim_to_interpolation_params(elev, azim):
    """(before, a, after) such that azimuth `azim` on elevation ring `elev` lies between grid rows
    `before` and `after` with weight `a` (sphere.py:78-121).  Raises AssertionError for a NaN
    azimuth (sphere.py:87) and ValueError for an elevation that is not a grid ring (:100-101)."""
    kind = az_kind(azim)
    before, after, alpha = C.c_int(), C.c_int(), C.c_double()
    rc = _cabi.lib.bas_ring_lookup_host(float(elev), float(azim), kind, C.byref(before), C.byref(alpha), C.byref(after))
    if rc == _cabi.ERR_AZIM_ASSERT:
        raise AssertionError('azim >= 0')
    if rc == _cabi.E_ARG:
        raise ValueError('ele must be one of the values in the database: [-45,-30,-15,0,15,30,45,60,75,90] .* (2pi / 360)')
    if rc != 0:
        raise _cabi.BasError(_cabi.last_error())
    if before.value == 186:
        return (186, 0., 186)
    a = alpha.value
    if kind == _cabi.AZ_F64:
        a = np.float64(a)
    else:
        a = np.float32(a)                           # exact: the weight was computed in float32
    return (before.value, a, after.value)
