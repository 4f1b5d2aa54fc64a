import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KIND_PY, KIND_F64, KIND_F32 = 0, 1, 2


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def as_kind(value, kind):
    return {KIND_PY: float, KIND_F64: np.float64, KIND_F32: np.float32}[int(kind)](value)


class GoldenBank:
    """Object with the five attributes of the reference's bank (apply_hrtf.py:36-44)."""

    def __init__(self, upsampling, diffs_left, diffs_right, irs_left, irs_right):
        self.upsampling = int(upsampling)
        self.diffs_left = np.asarray(diffs_left, dtype=np.float64)
        self.diffs_right = np.asarray(diffs_right, dtype=np.float64)
        self.irs_left = np.asarray(irs_left, dtype=np.float64)
        self.irs_right = np.asarray(irs_right, dtype=np.float64)


@pytest.fixture(scope='session')
def golden():
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.npz')
    return dict(np.load(path))


@pytest.fixture(scope='session')
def golden_bank(golden):
    return GoldenBank(golden['bank_upsampling'], golden['bank_diffs_left'], golden['bank_diffs_right'],
                      golden['bank_irs_left'], golden['bank_irs_right'])


@pytest.fixture(scope='session')
def oracle():
    from oracle import binaural_oracle
    return binaural_oracle


@pytest.fixture(scope='session')
def bas():
    import binaural_audio_synthesis_b200
    return binaural_audio_synthesis_b200


@pytest.fixture(scope='session')
def synth_bank(bas):
    """Full-size synthetic bank (U=8, K=256), like BASELINE.json configs 1-3."""
    f = bas.bank_synth.build_bank(8, seed=0)
    keep = 256 * 8
    return GoldenBank(8, f['diffs_left'], f['diffs_right'], f['irs_left'][:, :keep], f['irs_right'][:, :keep])


def golden_trajectory(name, k):
    """The trajectories make_golden.py used (apply_hrtf.py:583-588 with a short period)."""
    return {
        'circle': lambda t: (0, (k * t) % (2 * np.pi)),
        'askew': lambda t: ((np.pi / 4) * np.cos(k * t), (k * t) % (2 * np.pi)),
        'lissajous': lambda t: (np.deg2rad(22.5 + 67.5 * np.sin(3 * k * t + 0.3)), (5 * k * t + 1) % (2 * np.pi)),
        'passing': lambda t: (0, np.arctan(12 * np.cos(2 * k * t))),
    }[str(name)]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_abs_over_peak(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
