"""Generate tests/golden/reference_vectors.npz by running the UNMODIFIED reference
(/root/reference/apply_hrtf.py + sphere.py) in this container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs on a small synthetic bank are committed
as fixtures.  matplotlib is absent here and is imported (but never used on this path) at
apply_hrtf.py:17-18 / sphere.py:4-5, so empty stub modules are registered first.  The only
instrumentation is a recording wrapper around delay_signal_float that notes the delay argument of
every call (the integers floor/ceil of those delays are part of the parity contract); the wrapped
function itself is the reference's.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = '/root/reference'


def import_reference():
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import apply_hrtf
    import sphere
    return apply_hrtf, sphere


def golden_bank():
    """Small bank for the fixtures: the synthetic bank of bank_synth (seed 0, U=8) cut to K=32
    taps and rounded to float32 so it can be stored compactly; the delay tables stay float64."""
    sys.path.insert(0, ROOT)
    import binaural_audio_synthesis_b200 as bas
    fields = bas.bank_synth.build_bank(8, seed=0)
    keep = 32 * 8
    return dict(upsampling=8,
                diffs_left=fields['diffs_left'], diffs_right=fields['diffs_right'],
                irs_left=fields['irs_left'][:, :keep].astype(np.float32),
                irs_right=fields['irs_right'][:, :keep].astype(np.float32))


def bank_object(g, upsampling=None):
    class bank:
        pass
    bank.upsampling = int(upsampling or g['upsampling'])
    bank.diffs_left = np.asarray(g['diffs_left'], dtype=np.float64)
    bank.diffs_right = np.asarray(g['diffs_right'], dtype=np.float64)
    bank.irs_left = np.asarray(g['irs_left'], dtype=np.float64)
    bank.irs_right = np.asarray(g['irs_right'], dtype=np.float64)
    return bank


KIND_PY, KIND_F64, KIND_F32 = 0, 1, 2


def as_kind(v, kind):
    return {KIND_PY: float, KIND_F64: np.float64, KIND_F32: np.float32}[kind](v)


def direction_cases():
    """(elev, azim, kind) triples: grid points, wrap-around, sparse rings, pole, clamps, negative
    and > 2 pi azimuths, and seeded random directions, each in several scalar types."""
    d = np.deg2rad
    cases = []
    for kind in (KIND_PY, KIND_F64):
        for elev_deg, az_deg in [(0, 0), (0, 15), (0, 14.999), (0, 345), (0, 352.5), (0, 359.999), (0, 360), (0, 367),
                                 (0, -10), (0, 725), (10, 100), (-45, 200), (-50, 33), (-90, 1), (45, 7.5), (50, 171),
                                 (60, 29), (60, 331), (66, 45), (75, 59.9), (75, 300), (80, 123), (89.9, 10), (90, 77),
                                 (95, 200), (22.5, 180), (-15, 90), (30, 270), (37.2, 181.3), (-44.9, 0.01)]:
            cases.append((float(d(elev_deg)), float(d(az_deg)), kind))
    rng = np.random.default_rng(7)
    for i in range(40):
        cases.append((float(rng.uniform(-1.0, 1.7)), float(rng.uniform(-7, 14)), (KIND_PY, KIND_F64, KIND_F32)[i % 3]))
    # exact float32 grid azimuths in all three types (the SURVEY section-5 hazard)
    for kind in (KIND_PY, KIND_F64, KIND_F32):
        cases.append((0.0, float(np.float32(15) * np.float32(2 * np.pi / 360)), kind))
        cases.append((float(d(60)), float(np.float32(30) * np.float32(2 * np.pi / 360)), kind))
    return cases


def main():
    ref, sphere = import_reference()
    g = golden_bank()
    bank = bank_object(g)
    out = {'bank_' + k: v for k, v in g.items()}

    recorded = []
    original = ref.delay_signal_float

    def recording(in_sig, samples, downsample=1):
        recorded.append(float(samples))
        return original(in_sig, samples, downsample)

    ref.delay_signal_float = recording

    # 1. ring lookups (sphere.py:78-121)
    ring = []
    ring_elevs = np.deg2rad(np.array([-45, -30, -15, 0, 15, 30, 45, 60, 75, 90]))
    rng = np.random.default_rng(11)
    for e in ring_elevs:
        for kind in (KIND_PY, KIND_F64, KIND_F32):
            azs = list(rng.uniform(-7, 14, 6)) + [0.0, float(np.float32(2 * np.pi)), 2 * np.pi, 6.2831, 1e-9,
                                                   float(np.float32(45) * np.float32(2 * np.pi / 360))]
            for az in azs:
                b, a, aft = sphere.azim_to_interpolation_params(e, as_kind(az, kind))
                ring.append((e, az, kind, b, float(a), aft))
    out['ring_cases'] = np.array(ring, dtype=np.float64)

    # 2. interpolate_2d (apply_hrtf.py:171-281) with the twelve delays of every call
    cases = direction_cases()
    irs, delays = [], []
    for elev, azim, kind in cases:
        recorded.clear()
        irs.append(ref.interpolate_2d(bank, elev, as_kind(azim, kind)))
        assert len(recorded) == 12
        delays.append(list(recorded))
    out['dir_cases'] = np.array(cases, dtype=np.float64)
    out['dir_irs'] = np.array(irs)
    out['dir_delays'] = np.array(delays)

    # 3. ring interpolation (apply_hrtf.py:53-106), both return_upsampled modes
    ring_in, ring_out_dec, ring_out_up, ring_delays = [], [], [], []
    rng = np.random.default_rng(13)
    pairs = [(72, 73, 0.0), (72, 73, 1.0), (95, 72, 0.37), (186, 186, 0.0), (10, 11, 0.5), (170, 171, 0.999),
             (185, 180, 0.25), (0, 23, 0.6)] + [(int(rng.integers(0, 187)), int(rng.integers(0, 187)), float(rng.uniform())) for _ in range(8)]
    for b, a, alpha in pairs:
        dl, dr, dec = ref.delay_compensated_interpolation_with_delaydiff(bank, b, a, alpha)
        _, _, up = ref.delay_compensated_interpolation_with_delaydiff(bank, b, a, alpha, return_upsampled=True)
        ring_in.append((b, a, alpha)); ring_out_dec.append(dec); ring_out_up.append(up); ring_delays.append((dl, dr))
    out['ringinterp_in'] = np.array(ring_in)
    out['ringinterp_dec'] = np.array(ring_out_dec)
    out['ringinterp_up'] = np.array(ring_out_up)
    out['ringinterp_delays'] = np.array(ring_delays)

    # 4. make_signal_move_2d (apply_hrtf.py:356-466)
    ref.delay_signal_float = original
    fs = 44100
    k = 2 * np.pi / (0.05 * fs)          # fast motion so a short signal crosses many grid cells
    trajectories = {
        'circle': lambda t: (0, (k * t) % (2 * np.pi)),                                             # apply_hrtf.py:585
        'askew': lambda t: ((np.pi / 4) * np.cos(k * t), (k * t) % (2 * np.pi)),                    # :586
        'lissajous': lambda t: (np.deg2rad(22.5 + 67.5 * np.sin(3 * k * t + 0.3)), (5 * k * t + 1) % (2 * np.pi)),
        'passing': lambda t: (0, np.arctan(12 * np.cos(2 * k * t))),                                # :588
    }
    renders = [('circle', 3000, 512, 32, 0.05), ('lissajous', 2500, 512, 32, 0.05), ('askew', 1111, 256, 64, 0.05),
               ('passing', 2048, 512, 512, 0.05), ('lissajous', 1500, 512, 32, 3.0), ('circle', 700, 96, 32, 0.05)]
    rng = np.random.default_rng(17)
    for i, (name, n, c, s, sigma) in enumerate(renders):
        x = (sigma * rng.standard_normal(n)).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            y = ref.make_signal_move_2d(x, c, s, trajectories[name], bank)
        out['render%d_x' % i] = x
        out['render%d_y' % i] = np.ascontiguousarray(y)
        out['render%d_meta' % i] = np.array([n, c, s], dtype=np.int64)
        out['render%d_traj' % i] = np.array(name)
    out['render_k'] = np.array(k)
    out['n_renders'] = np.array(len(renders))

    path = os.path.join(HERE, 'reference_vectors.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
