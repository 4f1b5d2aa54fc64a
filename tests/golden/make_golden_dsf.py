"""Generate tests/golden/reference_dsf.npz: the reference's delay_signal_float (apply_hrtf.py:127-165)
run UNMODIFIED in this container on seeded signals: negative, integer, fractional and beyond-one-period
delays, with and without decimation, odd and even lengths.

    python tests/golden/make_golden_dsf.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference   # noqa: E402


def cases():
    """(length, delay, downsample)"""
    out = [(2048, 0.0, 1), (2048, 3.0, 1), (2048, -3.0, 1), (2048, 2.25, 1), (2048, -7.625, 1), (2048, 17.3, 8),
           (2048, -0.4, 8), (2048, 2047.5, 8), (2048, 5000.1, 1), (2048, -5000.9, 16), (8192, 33.333, 16), (8192, -12.0, 16),
           (1000, 7.77, 1), (1001, -2.5, 7), (37, 40.5, 1), (37, 0.999999, 3), (1, 0.5, 1), (2, -1.5, 2), (512, 1e-9, 1),
           (512, -1e-9, 4)]
    rng = np.random.default_rng(7)
    for _ in range(12):
        out.append((int(rng.integers(3, 1500)), float(rng.uniform(-300, 300)), int(rng.choice([1, 1, 2, 8, 16]))))
    return out


def main():
    ref, _ = import_reference()
    rng = np.random.default_rng(11)
    out = {'cases': np.array(cases(), dtype=np.float64)}
    for i, (n, d, ds) in enumerate(cases()):
        x = rng.standard_normal(n)
        out['x%d' % i] = x
        out['y%d' % i] = ref.delay_signal_float(x, d, ds)
    np.savez_compressed(os.path.join(HERE, 'reference_dsf.npz'), **out)
    print('wrote reference_dsf.npz: %d cases' % len(cases()))


if __name__ == '__main__':
    main()
