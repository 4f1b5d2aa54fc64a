"""Oracle against the LIVE unmodified reference, when /root/reference is mounted (build container
only; the GPU box has no reference and these tests skip there)."""
import contextlib
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.exists('/root/reference/apply_hrtf.py'), reason='reference not mounted')


@pytest.fixture(scope='module')
def ref():
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.import_reference()


def test_random_directions_and_types(ref, oracle, golden_bank):
    apply_hrtf, sphere = ref
    rng = np.random.default_rng(2024)
    for i in range(150):
        elev, azim = rng.uniform(-1.2, 1.9), rng.uniform(-9, 18)
        az = (float, np.float64, np.float32)[i % 3](azim)
        want = apply_hrtf.interpolate_2d(golden_bank, elev, az)
        got = oracle.interpolate_2d(golden_bank, elev, az)
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max())


def test_plan_integers_vs_live_reference(ref, bas, golden_bank):
    """Host plan (plan_math.h) rows/weights against live sphere.azim_to_interpolation_params on a
    dense sweep, including every exact float32 grid azimuth in all three scalar types."""
    _, sphere = ref
    grid = sphere.index_elev_azim
    checked = 0
    for kind, ctor in ((0, float), (1, np.float64), (2, np.float32)):
        for ring_elev in np.deg2rad([-45, -30, -15, 0, 15, 30, 45, 60, 75]):
            azs = list(np.unique(grid[:, 2]).astype(np.float64)) + list(np.linspace(-6.5, 13, 57))
            for az in azs:
                want = sphere.azim_to_interpolation_params(ring_elev, ctor(az))
                got = bas.sphere.azim_to_interpolation_params(ring_elev, ctor(az))
                assert (got[0], got[2]) == (want[0], want[2]) and float(got[1]) == float(want[1]), (kind, ring_elev, az)
                checked += 1
    assert checked > 2000


def test_render_matches_live_reference(ref, oracle, golden_bank):
    apply_hrtf, _ = ref
    rng = np.random.default_rng(5)
    x = (0.05 * rng.standard_normal(1800)).astype(np.float32)
    traj = lambda t: (0.5 * np.sin(0.002 * t), (0.005 * t) % (2 * np.pi))
    with contextlib.redirect_stdout(io.StringIO()):
        want = apply_hrtf.make_signal_move_2d(x, 256, 32, traj, golden_bank)
    got = oracle.make_signal_move_2d(x, 256, 32, traj, golden_bank)
    assert got.shape == want.shape
    assert np.linalg.norm(got - want) <= 2e-7 * np.linalg.norm(want)

