"""Trial vectorisation of plain trajectory callables (apply_hrtf.evaluate_trajectory): a callable the
reference would call 5,169 times with a Python int (apply_hrtf.py:429, :435) is called once with the
array of chunk boundaries - but only if that provably gives the reference's own values and scalar
kinds; otherwise the loop runs.  No GPU needed."""
import math

import numpy as np
import pytest

from .conftest import KIND_PY, KIND_F64, KIND_F32

FS = 44100
TIMES = np.arange(0, 60 * FS + 512, 512, dtype=np.int64)


def reference_loop(bas, fn, times):
    """What the reference does: one call per boundary with a Python int."""
    elev = np.empty(len(times))
    azim = np.empty(len(times))
    kinds = np.empty(len(times), dtype=np.uint8)
    for i, t in enumerate(times):
        e, a = fn(int(t))
        elev[i], azim[i], kinds[i] = e, a, bas.sphere.az_kind(a)
    return elev, azim, kinds


@pytest.mark.parametrize('name', ['circle_front', 'circle_horizontal', 'circle_askew', 'halfcircle_vertical', 'passing', 'spiral'])
def test_reference_main_trajectories_vectorise_exactly(bas, name):
    """All six lambdas of the reference's main (apply_hrtf.py:583-593) take the array path and give
    bit-identical directions and the same azimuth kind as the scalar loop."""
    from binaural_audio_synthesis_b200 import cli
    fn = cli.trajectories(FS)[name]
    calls = []

    def counted(t):
        calls.append(np.ndim(t))
        return fn(t)
    state = {}
    elev, azim, kind = bas.apply_hrtf.evaluate_trajectory(counted, TIMES, state)
    assert state['mode'] == 'array'
    assert len(calls) <= bas.apply_hrtf.TRAJECTORY_CHECKS + 2 and calls.count(1) == 1
    want_e, want_a, want_k = reference_loop(bas, fn, TIMES)
    assert np.array_equal(elev, want_e) and np.array_equal(azim, want_a)
    assert (want_k == kind).all()


def test_python_float_azimuth_keeps_its_kind(bas):
    """A trajectory whose scalar call returns a Python float selects the float32 ring arithmetic
    (SURVEY.md section 5); the array call's float64 array must not change that."""
    fn = lambda t: (0.0, (1e-4 * t) % 6.0)                # int * float -> Python float for an int t
    assert isinstance(fn(5)[1], float)
    _, _, kind = bas.apply_hrtf.evaluate_trajectory(fn, TIMES)
    assert kind == KIND_PY
    fn64 = lambda t: (0.0, np.float64(1e-4) * t)
    assert bas.apply_hrtf.evaluate_trajectory(fn64, TIMES)[2] == KIND_F64
    fn32 = lambda t: (0.0, np.float32(1e-4) * np.float32(t))
    e, a, kind = bas.apply_hrtf.evaluate_trajectory(fn32, TIMES)
    assert kind == KIND_F32 and np.array_equal(a, reference_loop(bas, fn32, TIMES)[1])


def test_non_vectorisable_callables_fall_back_to_the_loop(bas):
    cases = {
        'math module': lambda t: (0.0, math.fmod(1e-4 * t, 6.0)),                       # TypeError on arrays
        'branch on t': lambda t: (0.1, 1.0) if t < 30 * FS else (0.2, np.float64(2.0)),  # ambiguous truth value
        'table lookup': lambda t: (0.0, [0.5, 1.5, 2.5][t % 3]),                         # list index with an array
    }
    for name, fn in cases.items():
        state = {}
        elev, azim, kinds = bas.apply_hrtf.evaluate_trajectory(fn, TIMES, state)
        assert state['mode'] == 'loop', name
        want_e, want_a, want_k = reference_loop(bas, fn, TIMES)
        assert np.array_equal(elev, want_e) and np.array_equal(azim, want_a), name
        assert np.array_equal(np.broadcast_to(kinds, want_k.shape), want_k), name


def test_array_result_that_differs_from_scalar_calls_is_rejected(bas):
    """A callable that accepts arrays but computes something else for them must not be trusted."""
    def fn(t):
        if isinstance(t, np.ndarray):
            return np.zeros(t.shape), 1e-4 * t + 1e-9
        return 0.0, np.float64(1e-4 * t)
    state = {}
    elev, azim, kinds = bas.apply_hrtf.evaluate_trajectory(fn, TIMES, state)
    assert state['mode'] == 'loop'
    assert np.array_equal(azim, 1e-4 * TIMES)


def test_opt_out_and_opt_in(bas):
    fn = lambda t: (0.0, np.float64(1e-4) * t)
    fn.vectorized = False
    calls = []
    counted = lambda t: (calls.append(1), fn(t))[1]
    counted.vectorized = False
    bas.apply_hrtf.evaluate_trajectory(counted, TIMES[:100])
    assert len(calls) == 100
    declared = lambda t: (np.zeros(len(t)), np.float64(1e-4) * t)
    declared.vectorized = True
    e, a, kind = bas.apply_hrtf.evaluate_trajectory(declared, TIMES)
    assert kind == KIND_F64 and a.shape == TIMES.shape


def test_short_trajectories_just_loop(bas):
    calls = []
    fn = lambda t: (calls.append(np.ndim(t)), (0.0, 1e-4 * t))[1]
    bas.apply_hrtf.evaluate_trajectory(fn, TIMES[:8])
    assert calls == [0] * 8


def test_suggest_chunk_sizes_follows_the_source_speed(bas):
    """apply_hrtf.py:383-385, the reference's "idea for the future": faster sources get smaller chunks; every
    suggestion is a pair the tiled kernel takes (subchunksize 16 / 32 / 64 dividing the chunksize)."""
    from binaural_audio_synthesis_b200 import cli
    slow = cli.trajectories(FS, period=60.0)['circle_horizontal']
    usual = cli.trajectories(FS, period=4.0)['circle_horizontal']
    fast = cli.trajectories(FS, period=0.25)['circle_horizontal']
    sizes = [bas.suggest_chunk_sizes(f, 5 * FS) for f in (slow, usual, fast)]
    assert sizes[0][0] >= sizes[1][0] >= sizes[2][0] and sizes[0][1] >= sizes[1][1] >= sizes[2][1]
    assert sizes[2] == (128, 16) and sizes[0] == (1024, 64)
    for c, s in sizes:
        assert c % s == 0 and s in (16, 32, 64)
    # the fastest stretch decides, and per-chunk rotation stays below the bound
    c, s = bas.suggest_chunk_sizes(usual, 5 * FS, max_chunk_degrees=6.0)
    assert c * 360.0 / (4.0 * FS) <= 6.0
    assert bas.suggest_chunk_sizes(lambda t: (0.3, 1.0), 5 * FS) == (1024, 64)          # a source that does not move
