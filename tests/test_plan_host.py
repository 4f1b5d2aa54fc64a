"""Index parity of the plan arithmetic (plan_math.h, compiled for host AND device from one source)
against the reference's golden vectors - runs without a GPU through the C-ABI host entry points."""
import ctypes as C
import re

import numpy as np
import pytest

from .conftest import as_kind, KIND_PY, KIND_F64, KIND_F32


def eval_terms(bank, terms):
    """Evaluate the merged gather terms the way ir_synth.cu does, in float64 on the host."""
    u, length = bank.upsampling, bank.irs_left.shape[1]
    m = np.arange(0, length, u)
    out = []
    for ear, irs in enumerate((bank.irs_left, bank.irs_right)):
        acc = np.zeros(m.size)
        for rs, w in terms[ear]:
            if w != 0:
                acc += float(w) * irs[int(rs) >> 20, (m - (int(rs) & 0xFFFFF)) % length]
        out.append(acc)
    return np.vstack(out)


def test_header_symbols_exported(bas):
    """Every function include/bas_b200.h declares is exported by the library."""
    import os
    header = open(os.path.join(os.path.dirname(bas._cabi._HERE), 'include', 'bas_b200.h')).read()
    declared = set(re.findall(r'^(?:int|long long) (bas_\w+)\(', header, flags=re.M))
    assert len(declared) >= 13
    lib = C.CDLL(bas._cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(bas._cabi.EXPORTS)
    assert lib.bas_abi_version() == bas._cabi.ABI_VERSION == 5


def test_probe_library_is_separate(bas):
    """Measurement kernels live in libbas_probe.so (include/bas_probe.h); the product library exports none."""
    import os
    root = os.path.dirname(bas._cabi._HERE)
    header = open(os.path.join(root, 'include', 'bas_probe.h')).read()
    declared = set(re.findall(r'^int (bas_probe_\w+)\(', header, flags=re.M))
    assert declared
    probe = C.CDLL(os.path.join(bas._cabi._HERE, 'libbas_probe.so'))
    product = C.CDLL(bas._cabi.LIB_PATH)
    for name in declared:
        assert hasattr(probe, name), name
        assert not hasattr(product, name), name


def test_ring_lookup_host_bit_exact(bas, golden):
    for e, az, kind, b, a, aft in golden['ring_cases']:
        got = bas.sphere.azim_to_interpolation_params(e, as_kind(az, kind))
        assert (got[0], got[2]) == (int(b), int(aft)), (e, az, kind)
        assert float(got[1]) == a, (e, az, kind, got[1], a)


def test_ring_lookup_errors(bas):
    with pytest.raises(ValueError):
        bas.sphere.azim_to_interpolation_params(0.1, 1.0)            # not a grid ring: sphere.py:100-101
    with pytest.raises(AssertionError):
        bas.sphere.azim_to_interpolation_params(0.0, float('nan'))   # sphere.py:87
    assert bas.sphere.azim_to_interpolation_params(np.pi / 2, 2.0) == (186, 0., 186)   # sphere.py:92-93
    assert bas.sphere.azim_to_interpolation_params(2.0, 2.0) == (186, 0., 186)         # clipped to the pole (:88)
    assert bas.sphere.azim_to_interpolation_params(-2.0, 0.3)[0] == 1                  # clipped to -45 deg


def test_plan_integers_bit_exact(bas, golden, golden_bank):
    cases = golden['dir_cases']
    terms, trace = bas.plan_points_host(golden_bank, cases[:, 0], cases[:, 1], cases[:, 2].astype(np.uint8))
    assert not trace['err'].any()
    for i, delays in enumerate(golden['dir_delays']):
        lo = np.floor(delays).astype(np.int64)
        hi = np.ceil(delays).astype(np.int64)
        for ear in range(2):
            order = [0 + ear, 2 + ear, 4 + ear, 6 + ear, 8 + ear, 10 + ear]      # reference call order -> trace slots
            assert list(trace['lo'][i, ear]) == [lo[j] for j in order], (i, ear)
            assert list(trace['hi'][i, ear]) == [hi[j] for j in order], (i, ear)


def test_plan_rows_and_weights(bas, oracle, golden, golden_bank):
    for i, (elev, azim, kind) in enumerate(golden['dir_cases']):
        lower, higher = oracle.bracketing_rings(elev)
        tb, ta_w, ta = oracle.ring_neighbours(higher, as_kind(azim, kind))
        bb, ba_w, ba = oracle.ring_neighbours(lower, as_kind(azim, kind))
        _, trace = bas.plan_points_host(golden_bank, [elev], [azim], int(kind))
        assert list(trace['rows'][0]) == [tb, ta, bb, ba]
        assert trace['alpha_top'][0] == float(ta_w) and trace['alpha_bot'][0] == float(ba_w)


def test_merged_terms_reproduce_reference_irs(bas, golden, golden_bank):
    cases = golden['dir_cases']
    terms, _ = bas.plan_points_host(golden_bank, cases[:, 0], cases[:, 1], cases[:, 2].astype(np.uint8))
    for i, want in enumerate(golden['dir_irs']):
        got = eval_terms(golden_bank, terms[i])
        scale = max(1.0, np.abs(want).max())
        assert np.abs(got - want).max() <= 3e-7 * scale, i       # weights are stored as float32


def test_plan_error_bits(bas, golden_bank):
    _, trace = bas.plan_points_host(golden_bank, [0.0, float('nan'), 0.3], [float('nan'), 1.0, 1.0], KIND_F64)
    assert trace['err'][0] & bas._cabi.ERR_AZIM_ASSERT
    assert trace['err'][1] & bas._cabi.ERR_VERT_ASSERT
    assert trace['err'][2] == 0


def test_az_kind_classification(bas):
    k = bas.sphere.az_kind
    assert k(1.0) == KIND_PY and k(1) == KIND_PY
    assert k(np.float64(1)) == KIND_F64 and k(np.float32(1)) == KIND_F32
    assert k(np.int64(1)) == KIND_F64 and k(np.array(1.0)) == KIND_F64


def test_no_cpu_fallback(bas, golden_bank):
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(bas.BasError):
        bas.interpolate_2d(golden_bank, 0.1, 0.2)
    with pytest.raises(bas.BasError):
        bas.make_signal_move_2d(np.zeros(1024, dtype=np.float32), 512, 32, lambda t: (0, 0.0), golden_bank)
