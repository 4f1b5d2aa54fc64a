"""Generate tests/golden/reference_legacy.npz: the reference's legacy 1-D renderer make_signal_move
(apply_hrtf.py:294-353) and delay_compensated_interpolation_easy (:114-125), run UNMODIFIED in this
container on the golden bank of make_golden.py.

    python tests/golden/make_golden_legacy.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, golden_bank, bank_object   # noqa: E402


def index_functions(period):
    """Continuous indices on the horizontal ring (rows 72..96, 97 wraps to 73)."""
    return {
        'sweep': lambda t: 72 + 24.999 * ((t / period) % 1.0),
        'wobble': lambda t: 84 + 10 * np.sin(2 * np.pi * t / period),
        'fixed': lambda t: 80.25,
        'integer': lambda t: 72 + (t // 64) % 25,
    }


def main():
    ref, _ = import_reference()
    g = golden_bank()
    bank = bank_object(g)
    out = {}
    easy = [72.0, 72.5, 80.25, 95.999, 96.0, 96.3, 73.0, 85.0]
    out['easy_in'] = np.array(easy)
    out['easy_out'] = np.array([ref.delay_compensated_interpolation_easy(bank, v) for v in easy])
    rng = np.random.default_rng(23)
    cases = [('sweep', 64, 1000, 0.1), ('wobble', 128, 777, 0.1), ('fixed', 32, 300, 0.1), ('integer', 64, 640, 0.1), ('sweep', 96, 500, 30.0)]
    for i, (name, chunk, n, scale) in enumerate(cases):
        x = (scale * rng.standard_normal(n)).astype(np.float32)
        fn = index_functions(400.0)[name]
        with contextlib.redirect_stdout(io.StringIO()):
            y = ref.make_signal_move(x, chunk, fn, bank)
        out['legacy%d_x' % i] = x
        out['legacy%d_y' % i] = y
        out['legacy%d_meta' % i] = np.array([name, str(chunk)])
    np.savez_compressed(os.path.join(HERE, 'reference_legacy.npz'), **out)
    print('wrote reference_legacy.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
