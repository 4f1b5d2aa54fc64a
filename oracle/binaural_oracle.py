"""CPU oracle for the moving-source binaural renderer.  TEST INFRASTRUCTURE ONLY.

This module is a float64 numpy restatement of the hot path of the reference
(`/root/reference/apply_hrtf.py` + `/root/reference/sphere.py`).  It is the
checker the CUDA path is compared against; it is never imported by the product
package (`binaural-audio-synthesis_b200/`).  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against the *live, unmodified reference* imported in the build
container (`tests/golden/make_golden.py` writes `tests/golden/*.npz`;
`tests/test_reference_live.py` re-checks live when `/root/reference` exists).

Every function cites the reference lines it restates.  The arithmetic (operation
order, scalar dtypes, NumPy-2 weak-scalar promotion) follows the reference so that
all integer quantities (grid rows, floor/ceil delays) are bit-identical and the
float64 values agree to rounding.
"""
from __future__ import annotations

import math
import numpy as np

TWO_PI = 2 * np.pi

# --------------------------------------------------------------------------------------
# sphere grid  (sphere.py:124-319, :350)
# --------------------------------------------------------------------------------------
# The measurement grid: rings at -45..45 deg hold 24 points (15 deg apart), 60 deg holds 12
# (30 deg apart), 75 deg holds 6 (60 deg apart), 90 deg is the single pole point.
RING_ELEV_DEG = (-45, -30, -15, 0, 15, 30, 45, 60, 75, 90)
RING_COUNT = (24, 24, 24, 24, 24, 24, 24, 12, 6, 1)
RING_START = tuple(int(v) for v in np.concatenate([[0], np.cumsum(RING_COUNT)[:-1]]))
N_DIRECTIONS = 187


def grid_table() -> np.ndarray:
    """(187, 3) float32 table of (index, elev_rad, azim_rad).

    sphere.py:127-315 lists the rows literally as float32; sphere.py:318 converts the two
    angle columns to radians with an in-place float32 multiply by the Python float
    2*pi/360 (i.e. by float32(2*pi/360)).  Generated here from the ring structure.
    """
    rows = []
    for elev, count, start in zip(RING_ELEV_DEG, RING_COUNT, RING_START):
        step = 360 // count if count > 1 else 0
        for k in range(count):
            rows.append((start + k, elev, k * step))
    tab = np.array(rows, dtype=np.float32)
    tab[:, 1:3] *= (2 * np.pi / 360)
    return tab


GRID = grid_table()


def ring_neighbours(ring_elev, azim):
    """(before, weight, after) on one elevation ring.  Restates sphere.py:78-121.

    `azim` keeps whatever scalar type the caller passed: a Python float is a weak scalar,
    so the comparisons against the float32 table and the weight are evaluated in float32;
    an np.float64 forces float64 (SURVEY.md section 5 'dtype hazard').
    """
    azim = azim % TWO_PI                                           # sphere.py:86
    if not (azim >= 0):                                            # sphere.py:87
        raise AssertionError('azimuth is negative or NaN after the modulo')
    ring_elev = np.clip(ring_elev, -np.pi / 4, np.pi / 2)          # sphere.py:88
    tol = 0.00001                                                  # sphere.py:90
    if abs(ring_elev - np.pi / 2) < tol:                           # sphere.py:92-93
        return (186, 0., 186)
    on_ring = np.flatnonzero(np.abs(GRID[:, 1] - ring_elev) < tol)  # sphere.py:98
    if on_ring.size == 0:                                          # sphere.py:100-101
        raise ValueError('elevation is not one of the grid rings')
    ring_az = GRID[on_ring, 2]
    before = int(on_ring[ring_az <= azim].max())                   # sphere.py:103
    later = on_ring[ring_az > azim]
    after = int(later.min()) if later.size else int(on_ring[0])   # sphere.py:104-109
    az_before = GRID[before, 2]                                    # sphere.py:112-113
    az_after = GRID[after, 2]
    if az_after < az_before:                                       # sphere.py:115-117
        assert az_after == 0
        az_after = 2 * np.pi
    weight = (azim - az_before) / (az_after - az_before)           # sphere.py:119
    return (before, weight, after)


# --------------------------------------------------------------------------------------
# bank type  (apply_hrtf.py:23-46)
# --------------------------------------------------------------------------------------
class Bank:
    """Plain holder with the five attributes of the reference's class-as-struct
    (apply_hrtf.py:36-44)."""

    def __init__(self, upsampling, diffs_left, diffs_right, irs_left, irs_right):
        self.upsampling = int(upsampling)
        self.diffs_left = diffs_left
        self.diffs_right = diffs_right
        self.irs_left = irs_left
        self.irs_right = irs_right


def load_bank(filename, samples_to_keep=512) -> Bank:
    """apply_hrtf.py:23-46: read the .mat struct, keep samples_to_keep*U samples per row."""
    import scipy.io
    m = scipy.io.loadmat(filename)['irs_and_delaydiffs'][0][0]
    u = int(m['upsampling'][0][0])
    return Bank(u, m['diffs_left'], m['diffs_right'],
                m['irs_left'][:, :samples_to_keep * u], m['irs_right'][:, :samples_to_keep * u])


# --------------------------------------------------------------------------------------
# fractional delay  (apply_hrtf.py:127-165)
# --------------------------------------------------------------------------------------
def split_delay(d):
    """(floor, ceil, frac) of a delay exactly as apply_hrtf.py:149-151."""
    lo = int(np.floor(d))
    hi = int(np.ceil(d))
    return lo, hi, d - lo


def circular_delay(sig: np.ndarray, d, step: int = 1) -> np.ndarray:
    """Two-tap linear-interpolated *circular* delay, optionally decimated by `step`.

    apply_hrtf.py:156-157 uses np.roll (wrap-around, no zero fill); :160-163 keeps every
    step-th sample of the rolled signals; :165 blends.
    """
    lo, hi, frac = split_delay(d)
    n = np.arange(0, sig.size, step) if step > 1 else np.arange(sig.size)
    size = sig.size
    return (1 - frac) * sig[(n - lo) % size] + frac * sig[(n - hi) % size]


def delay_signal_float(in_sig, samples, downsample=1) -> np.ndarray:
    """apply_hrtf.py:127-165 under the reference's own name and signature (circular_delay above)."""
    return circular_delay(np.asarray(in_sig), samples, int(downsample))


# --------------------------------------------------------------------------------------
# ring (1-D) interpolation  (apply_hrtf.py:53-106)
# --------------------------------------------------------------------------------------
def ring_interpolation(bank, before: int, after: int, alpha, return_upsampled=False, trace=None):
    """delay_compensated_interpolation_with_delaydiff, apply_hrtf.py:53-106."""
    u = bank.upsampling
    out = []
    deltas = []
    for ear, (diffs, irs) in enumerate(((bank.diffs_left, bank.irs_left),
                                        (bank.diffs_right, bank.irs_right))):
        d = u * diffs[before, after]                                 # :82-83
        aligned = circular_delay(irs[after, :], -d)                  # :86-87
        blend = (1 - alpha) * irs[before, :] + alpha * aligned       # :90-91
        d_back = alpha * d                                           # :94-95
        out.append(circular_delay(blend, d_back, 1 if return_upsampled else u))  # :97-102
        deltas.append(d_back / u)                                    # :106
        if trace is not None:
            trace.append(dict(ear=ear, remove=split_delay(-d)[:2], restore=split_delay(d_back)[:2]))
    return (deltas[0], deltas[1], np.vstack(out))                    # :104-106


# --------------------------------------------------------------------------------------
# 2-D interpolation  (apply_hrtf.py:171-281)
# --------------------------------------------------------------------------------------
_RING_ELEVS = np.deg2rad(np.array(RING_ELEV_DEG))                    # :199


def bracketing_rings(elev):
    """apply_hrtf.py:201-211: nearest grid ring at or below / at or above `elev`."""
    below = [e for e in _RING_ELEVS if e <= elev]
    lower = max(below) if below else -0.78539816339744828            # :201-204
    above = [e for e in _RING_ELEVS if e >= elev]
    higher = min(above) if above else 1.5707963267948966             # :206-209
    assert higher >= lower                                           # :211
    return lower, higher


def interpolate_2d(bank, elev, azim, trace=None):
    """apply_hrtf.py:171-281.  Returns (2, K) float64.

    If `trace` is a dict it receives every integer the parity contract covers: the four
    grid rows and, per ear, the floor/ceil pairs of the six fractional delays.
    """
    lower, higher = bracketing_rings(elev)
    tb, t_alpha, ta = ring_neighbours(higher, azim)                  # :214
    bb, b_alpha, ba = ring_neighbours(lower, azim)                   # :215
    t_trace, b_trace = [], []
    dlt, drt, top = ring_interpolation(bank, tb, ta, t_alpha, True, t_trace)   # :219
    dlb, drb, bot = ring_interpolation(bank, bb, ba, b_alpha, True, b_trace)   # :220
    u = bank.upsampling
    dv = (u * (-dlt + bank.diffs_left[tb, bb] + dlb),                # :246-248
          u * (-drt + bank.diffs_right[tb, bb] + drb))               # :250-252
    if higher > lower:                                               # :261-265
        a = (elev - lower) / (higher - lower)
    else:
        assert higher == lower
        a = 0
    assert 0 <= a <= 1                                               # :266
    ears = []
    for e in range(2):
        bot_aligned = circular_delay(bot[e, :], -dv[e])              # :254-255
        blend = (1 - a) * bot_aligned + a * top[e, :]                # :268-269
        ears.append(circular_delay(blend, (1 - a) * dv[e], u))       # :272-277
    if trace is not None:
        trace.update(rows=(tb, ta, bb, ba), alpha_top=t_alpha, alpha_bot=b_alpha, a=a,
                     top=t_trace, bot=b_trace,
                     vert_remove=[split_delay(-dv[e])[:2] for e in range(2)],
                     vert_restore=[split_delay((1 - a) * dv[e])[:2] for e in range(2)])
    return np.vstack(ears)                                           # :279-281


# --------------------------------------------------------------------------------------
# renderer  (apply_hrtf.py:356-466)
# --------------------------------------------------------------------------------------
def render_geometry(n_samples: int, chunk: int, sub: int, bank):
    """Lengths used by the renderer: apply_hrtf.py:399 (K), :401-402 (S | C), :405 (N_in),
    :410 (N_out)."""
    k = int(0.5 + bank.irs_left.shape[1] / bank.upsampling)
    ratio = chunk / sub
    assert ratio == np.floor(ratio), 'subchunksize does not divide chunksize evenly'
    n_in = int(0.5 + np.ceil(n_samples / chunk) * chunk)
    return k, n_in, n_in + k - 1


def boundary_filters(bank, trajectory, n_in: int, chunk: int) -> np.ndarray:
    """(n_in/chunk + 1, 2, K) float64: interpolate_2d at t = 0, C, ..., N_in inclusive
    (apply_hrtf.py:429, :435 - the last evaluation is past the end of the signal)."""
    return np.stack([interpolate_2d(bank, *trajectory(t)) for t in range(0, n_in + 1, chunk)])


def render_unnormalised(signal, chunk: int, sub: int, filters: np.ndarray, k: int) -> np.ndarray:
    """The chunk / subchunk overlap-add loops of apply_hrtf.py:431-453 in float64, returning
    planar (2, N_out).  Per subchunk: blend the two boundary filters with alpha = j/C (:442-443),
    full linear convolution of the S input samples with both ears (:445-446, direct method),
    add into the output at the subchunk's offset (:450-453)."""
    n_in = (filters.shape[0] - 1) * chunk
    x = np.zeros(n_in)
    x[:signal.size] = signal                                          # :406
    out = np.zeros((2, n_in + k - 1))                                 # :413-414
    for ci, i in enumerate(range(0, n_in, chunk)):
        h0, h1 = filters[ci], filters[ci + 1]                         # :434-435
        for j in range(0, chunk, sub):
            alpha = j / chunk                                         # :442
            h = (1 - alpha) * h0 + alpha * h1                         # :443
            seg = x[i + j:i + j + sub]
            out[0, i + j:i + j + sub + k - 1] += np.convolve(seg, h[0])   # :445, :452
            out[1, i + j:i + j + sub + k - 1] += np.convolve(seg, h[1])   # :446, :453
    return out


def finish(out_planar: np.ndarray) -> np.ndarray:
    """apply_hrtf.py:459-464: cast to float32, transpose to (N_out, 2), divide by the peak
    when it exceeds 1."""
    sig = out_planar.astype(np.float32).T
    peak = np.max([sig.max(), -(sig.min())])
    if peak > 1:
        sig /= peak
    return sig


def make_signal_move_2d(in_signal, chunksize: int, subchunksize: int, elev_azim_function, bank):
    """apply_hrtf.py:356-466 end to end (progress prints omitted)."""
    assert len(in_signal.shape) == 1, 'only mono signals for now'    # :398
    k, n_in, _ = render_geometry(in_signal.size, chunksize, subchunksize, bank)
    filters = boundary_filters(bank, elev_azim_function, n_in, chunksize)
    return finish(render_unnormalised(in_signal, chunksize, subchunksize, filters, k))


def ring_interpolation_easy(bank, continuous_index):
    """delay_compensated_interpolation_easy, apply_hrtf.py:114-125 (horizontal ring, 97 wraps to 73)."""
    before = int(np.floor(continuous_index))
    after = int(np.ceil(continuous_index))
    alpha = continuous_index - before
    if after == 97:
        after = 73
    return ring_interpolation(bank, before, after, alpha)[2]


def make_signal_move(in_signal, chunksize: int, index_function, bank):
    """The legacy 1-D renderer, apply_hrtf.py:294-353: one ring-interpolated filter per chunk (no
    cross-fade), full convolution of every chunk, overlap-add, float32 cast and peak division."""
    assert len(in_signal.shape) == 1, 'only mono signals for now'    # :308
    k = int(0.5 + bank.irs_left.shape[1] / bank.upsampling)         # :309
    n_in = int(0.5 + np.ceil(in_signal.size / chunksize) * chunksize)
    x = np.pad(np.asarray(in_signal), (0, n_in - in_signal.size), mode='constant')
    out = np.zeros((2, n_in + k - 1))
    for i in range(0, n_in, chunksize):                              # :331
        ir = ring_interpolation_easy(bank, index_function(i))       # :334
        for ear in range(2):
            out[ear, i:i + chunksize + k - 1] += np.convolve(x[i:i + chunksize], ir[ear])    # :337-343
    return finish(out)                                               # :348-353


# --------------------------------------------------------------------------------------
# closed forms used by property tests (derived in SURVEY.md section 3.2 / 3.3)
# --------------------------------------------------------------------------------------
def gather_terms(bank, elev, azim):
    """The 36 (row, shift, weight) gather terms per ear that interpolate_2d composes to:
    out_e[m] = sum_t w_t * bank_e[row_t][(m*U - shift_t) mod L].  Built from the same scalar
    arithmetic as interpolate_2d; used to check the merged-term plan of the CUDA path."""
    lower, higher = bracketing_rings(elev)
    tb, t_alpha, ta = ring_neighbours(higher, azim)
    bb, b_alpha, ba = ring_neighbours(lower, azim)
    u = bank.upsampling
    a = (elev - lower) / (higher - lower) if higher > lower else 0
    per_ear = []
    for diffs in (bank.diffs_left, bank.diffs_right):
        def ring_terms(before, after, alpha):
            d = u * diffs[before, after]
            lo1, hi1, f1 = split_delay(-d)
            lo2, hi2, f2 = split_delay(alpha * d)
            terms = []
            for s2, w2 in ((lo2, 1 - f2), (hi2, f2)):
                terms.append((before, s2, w2 * (1 - alpha)))
                for s1, w1 in ((lo1, 1 - f1), (hi1, f1)):
                    terms.append((after, s2 + s1, w2 * alpha * w1))
            return terms, (alpha * d) / u
        top_terms, d_top = ring_terms(tb, ta, t_alpha)
        bot_terms, d_bot = ring_terms(bb, ba, b_alpha)
        dv = u * (-d_top + diffs[tb, bb] + d_bot)
        lo3, hi3, f3 = split_delay(-dv)
        lo4, hi4, f4 = split_delay((1 - a) * dv)
        terms = []
        for s4, w4 in ((lo4, 1 - f4), (hi4, f4)):
            for (row, s, w) in top_terms:
                terms.append((row, s4 + s, w4 * a * w))
            for s3, w3 in ((lo3, 1 - f3), (hi3, f3)):
                for (row, s, w) in bot_terms:
                    terms.append((row, s4 + s3 + s, w4 * (1 - a) * w3 * w))
        per_ear.append(terms)
    return per_ear


def eval_gather_terms(bank, per_ear_terms):
    u = bank.upsampling
    size = bank.irs_left.shape[1]
    m = np.arange(0, size, u)
    out = []
    for irs, terms in zip((bank.irs_left, bank.irs_right), per_ear_terms):
        acc = np.zeros(m.size)
        for row, shift, w in terms:
            acc += w * irs[row, (m - shift) % size]
        out.append(acc)
    return np.vstack(out)


def render_closed_form(signal, chunk: int, sub: int, filters: np.ndarray, k: int) -> np.ndarray:
    """out_e[p] = sum_k x[p-k] * h_{q(p-k),e}[k]: the scatter form of the overlap-add loops
    (SURVEY.md section 3.2).  Vectorised over taps; used to cross-check render_unnormalised."""
    n_in = (filters.shape[0] - 1) * chunk
    x = np.zeros(n_in)
    x[:signal.size] = signal
    n = np.arange(n_in)
    ci = n // chunk
    alpha = ((n % chunk) // sub * sub) / chunk
    out = np.zeros((2, n_in + k - 1))
    for tap in range(k):
        for e in range(2):
            h = (1 - alpha) * filters[ci, e, tap] + alpha * filters[ci + 1, e, tap]
            out[e, tap:tap + n_in] += x * h
    return out
