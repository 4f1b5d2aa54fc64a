"""The CPU oracle (oracle/binaural_oracle.py) against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; everything on the GPU is then
compared with the oracle."""
import numpy as np

from .conftest import as_kind, golden_trajectory, rel_l2


def test_grid_table_matches_reference_layout(oracle):
    tab = oracle.GRID
    assert tab.shape == (187, 3) and tab.dtype == np.float32
    assert np.array_equal(tab[:, 0], np.arange(187, dtype=np.float32))
    # spot values of sphere.py:127-315 after the float32 radian conversion of :318
    f = np.float32(2 * np.pi / 360)
    assert tab[73, 2] == np.float32(15) * f and tab[72, 1] == 0
    assert tab[168, 1] == np.float32(60) * f and tab[169, 2] == np.float32(30) * f
    assert tab[185, 2] == np.float32(300) * f and tab[186, 1] == np.float32(90) * f


def test_ring_lookup_bit_exact(oracle, golden):
    for e, az, kind, b, a, aft in golden['ring_cases']:
        got = oracle.ring_neighbours(e, as_kind(az, kind))
        assert (got[0], got[2]) == (int(b), int(aft)), (e, az, kind)
        assert float(got[1]) == a, (e, az, kind)


def test_interpolate_2d_matches_reference(oracle, golden, golden_bank):
    for (elev, azim, kind), want, delays in zip(golden['dir_cases'], golden['dir_irs'], golden['dir_delays']):
        trace = {}
        got = oracle.interpolate_2d(golden_bank, elev, as_kind(azim, kind), trace)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
        # the integer floor/ceil of the twelve delays, in the reference's call order
        want_int = [(int(np.floor(d)), int(np.ceil(d))) for d in delays]
        got_int = ([trace['top'][0]['remove'], trace['top'][1]['remove'], trace['top'][0]['restore'], trace['top'][1]['restore'],
                    trace['bot'][0]['remove'], trace['bot'][1]['remove'], trace['bot'][0]['restore'], trace['bot'][1]['restore']]
                   + list(trace['vert_remove']) + list(trace['vert_restore']))
        assert [tuple(t) for t in got_int] == want_int


def test_gather_terms_closed_form(oracle, golden, golden_bank):
    """The 36-term gather the CUDA plan is built on reproduces interpolate_2d."""
    for (elev, azim, kind), want in list(zip(golden['dir_cases'], golden['dir_irs']))[::3]:
        terms = oracle.gather_terms(golden_bank, elev, as_kind(azim, kind))
        got = oracle.eval_gather_terms(golden_bank, terms)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_ring_interpolation_matches_reference(oracle, golden, golden_bank):
    for (b, a, alpha), dec, up, (dl, dr) in zip(golden['ringinterp_in'], golden['ringinterp_dec'],
                                                golden['ringinterp_up'], golden['ringinterp_delays']):
        got = oracle.ring_interpolation(golden_bank, int(b), int(a), float(alpha))
        assert got[0] == dl and got[1] == dr
        assert np.abs(got[2] - dec).max() <= 1e-13
        got_up = oracle.ring_interpolation(golden_bank, int(b), int(a), float(alpha), True)
        assert np.abs(got_up[2] - up).max() <= 1e-13


def test_render_matches_reference(oracle, golden, golden_bank):
    k = float(golden['render_k'])
    for i in range(int(golden['n_renders'])):
        n, c, s = (int(v) for v in golden['render%d_meta' % i])
        traj = golden_trajectory(golden['render%d_traj' % i], k)
        got = oracle.make_signal_move_2d(golden['render%d_x' % i], c, s, traj, golden_bank)
        want = golden['render%d_y' % i]
        assert got.shape == want.shape and got.dtype == np.float32
        assert rel_l2(got, want) <= 2e-7          # both are float32 casts of float64 sums


def test_render_closed_form_equals_overlap_add(oracle, golden, golden_bank):
    k = float(golden['render_k'])
    x = golden['render0_x']
    n, c, s = (int(v) for v in golden['render0_meta'])
    kk, n_in, _ = oracle.render_geometry(n, c, s, golden_bank)
    filters = oracle.boundary_filters(golden_bank, golden_trajectory('circle', k), n_in, c)
    a = oracle.render_unnormalised(x, c, s, filters, kk)
    b = oracle.render_closed_form(x, c, s, filters, kk)
    assert np.abs(a - b).max() <= 1e-14


def test_normalisation_branch(oracle, golden, golden_bank):
    """render4 was generated with sigma = 3 so the peak exceeds 1 (apply_hrtf.py:462-464)."""
    y = golden['render4_y']
    assert abs(np.abs(y).max() - 1.0) < 1e-6
