"""ctypes binding of libbas_b200.so (include/bas_b200.h).

There is no CPU fallback: if the library is missing the import fails with the build command, and
every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbas_b200.so')

ABI_VERSION = 5
N_DIRECTIONS = 187
MAX_TERMS = 16
AZ_PYFLOAT, AZ_F64, AZ_F32 = 0, 1, 2
ERR_AZIM_ASSERT, ERR_VERT_ASSERT, ERR_NONFINITE = 1, 2, 4
E_ARG, E_UNSUPPORTED, E_NO_DEVICE = -1, -2, -3
RENDER_AUTO, RENDER_GENERIC, RENDER_TILED = 0, 1, 2
RENDER_SPLIT, RENDER_NO_SPLIT = 0x40, 0x80
MIX_ACCUMULATE = 2
IR_UPSAMPLED, IR_PLANAR, IR_ROWS = 0, 1, 2
TILED_SHAPES = ((4, 2, 2), (4, 1, 2), (4, 1, 3), (6, 1, 2), (6, 2, 1), (8, 2, 1), (8, 1, 1))   # (warps per CTA, stages, CTAs per SM)


def render_variant(tw=0, ns=0, ctas=0, parts=0, split=None, base=RENDER_TILED):
    """bas_render variant word that requests a tile shape of the tiled kernel (0 = chosen by the library):
    tw warps per CTA, ns pipeline stages, ctas resident CTAs per SM, parts (1, 2, 4, 8) warps per
    1024-output stripe; split True / False forces / forbids splitting tiles between CTAs."""
    code = {0: 0, 1: 1, 2: 2, 4: 3, 8: 4}[parts]
    v = base | (tw << 8) | (ns << 16) | (ctas << 24) | (code << 28)
    if split is not None:
        v |= RENDER_SPLIT if split else RENDER_NO_SPLIT
    return v


class Term(C.Structure):
    _fields_ = [('row_shift', C.c_int32), ('weight', C.c_float)]


class Trace(C.Structure):
    _fields_ = [('rows', C.c_int32 * 4), ('err', C.c_int32), ('pad', C.c_int32),
                ('alpha_top', C.c_double), ('alpha_bot', C.c_double), ('a', C.c_double),
                ('lo', (C.c_int64 * 6) * 2), ('hi', (C.c_int64 * 6) * 2)]


TRACE_DTYPE = [('rows', '<i4', (4,)), ('err', '<i4'), ('pad', '<i4'), ('alpha_top', '<f8'),
               ('alpha_bot', '<f8'), ('a', '<f8'), ('lo', '<i8', (2, 6)), ('hi', '<i8', (2, 6))]
TERM_DTYPE = [('row_shift', '<i4'), ('weight', '<f4')]


class PipelineJob(C.Structure):
    """bas_pipeline_job (include/bas_b200.h)."""
    _fields_ = [('n_src', C.c_int), ('C', C.c_int), ('S', C.c_int), ('K', C.c_int), ('U', C.c_int), ('mix', C.c_int),
                ('variant', C.c_int), ('az_kind_all', C.c_int),
                ('n', C.c_longlong), ('n_in', C.c_longlong), ('p_begin', C.c_longlong), ('p_count', C.c_longlong),
                ('x_host_stride', C.c_longlong), ('segment_bytes', C.c_longlong),
                ('x_host', C.c_void_p), ('x_dev', C.c_void_p), ('dirs_host', C.c_void_p), ('az_kind_host', C.c_void_p),
                ('diffs_left_dev', C.c_void_p), ('diffs_right_dev', C.c_void_p), ('bank_pp_dev', C.c_void_p),
                ('out_host', C.c_void_p), ('small_host', C.c_void_p),
                ('arena_dev', C.c_void_p), ('arena_bytes', C.c_longlong),
                ('workspace_dev', C.c_void_p), ('workspace_bytes', C.c_longlong),
                ('stream_main', C.c_void_p), ('stream_up', C.c_void_p), ('stream_down', C.c_void_p),
                ('bank_pp2_dev', C.c_void_p)]


STEP_PLAN, STEP_RENDER, STEP_NORMALISE, STEP_FUSED = 1, 2, 4, 8


class Route(C.Structure):
    """bas_route (include/bas_b200.h)."""
    _fields_ = [('table_dev', C.c_void_p), ('n', C.c_int), ('rank', C.c_int), ('len', C.c_longlong), ('stride', C.c_longlong),
                ('arrive_ptrs_dev', C.c_void_p), ('arrive_counter_dev', C.c_void_p), ('arrive_epoch', C.c_uint), ('reserved', C.c_uint)]


class StepJob(C.Structure):
    """bas_step_job (include/bas_b200.h)."""
    _fields_ = [('n_src', C.c_int), ('C', C.c_int), ('S', C.c_int), ('K', C.c_int), ('U', C.c_int), ('mix', C.c_int),
                ('variant', C.c_int), ('az_kind_all', C.c_int), ('flags', C.c_int), ('reserved', C.c_int),
                ('n_valid', C.c_longlong), ('n_in', C.c_longlong), ('x_stride', C.c_longlong),
                ('p_begin', C.c_longlong), ('p_count', C.c_longlong), ('out_stride', C.c_longlong),
                ('x_dev', C.c_void_p), ('elev_dev', C.c_void_p), ('azim_dev', C.c_void_p), ('az_kind_dev', C.c_void_p),
                ('diffs_left_dev', C.c_void_p), ('diffs_right_dev', C.c_void_p), ('bank_pp_dev', C.c_void_p),
                ('bank_pp2_dev', C.c_void_p), ('terms_dev', C.c_void_p), ('filt_dev', C.c_void_p), ('gains_dev', C.c_void_p),
                ('out_dev', C.c_void_p), ('small_dev', C.c_void_p), ('workspace_dev', C.c_void_p),
                ('workspace_bytes', C.c_longlong), ('route', C.POINTER(Route))]


class BasError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libbas_b200.so is not built (%s).  Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C binaural-audio-synthesis_b200/csrc`.  There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, ll, i, dp = C.c_void_p, C.c_longlong, C.c_int, C.POINTER(C.c_double)
    sigs = {
        'bas_abi_version': ([], i),
        'bas_last_error': ([C.c_char_p, C.c_size_t], i),
        'bas_device_count': ([], i),
        'bas_bank_to_polyphase': ([vp, i, i, i, vp, vp], i),
        'bas_bank_upsample': ([vp, i, i, i, vp, i, vp, vp], i),
        'bas_bank_delay_diffs': ([vp, i, i, i, vp, i, vp, vp], i),
        'bas_plan_build': ([vp, vp, i, i, vp, vp, vp, i, ll, vp, vp, vp, vp], i),
        'bas_plan_build_host': ([vp, vp, i, i, vp, vp, vp, i, ll, vp, vp], i),
        'bas_ring_lookup_host': ([C.c_double, C.c_double, i, C.POINTER(i), dp, C.POINTER(i)], i),
        'bas_plan_ring': ([vp, vp, i, i, vp, vp, ll, vp, vp, vp, vp], i),
        'bas_plan_ring_host': ([vp, vp, i, i, i, i, C.c_double, C.c_double, vp, vp, vp, vp], i),
        'bas_delay_signal_float': ([vp, ll, ll, ll, C.c_double, i, vp, vp], i),
        'bas_ir_synth': ([vp, i, i, vp, ll, i, vp, ll, vp], i),
        'bas_filter_row_pitch': ([i], i),
        'bas_render': ([vp, ll, ll, i, ll, i, i, i, vp, vp, ll, ll, vp, ll, i, vp, i, vp, ll, vp], i),
        'bas_render_fused': ([vp, ll, ll, i, ll, i, i, i, vp, vp, i, vp, ll, ll, vp, ll, i, vp, i, vp, ll, vp], i),
        'bas_render_fused_supported': ([i, i], i),
        'bas_render_fused_shape': ([i], i),
        'bas_render_fused_fits': ([i, i, i, i, i], i),
        'bas_bank2_floats': ([i, i], ll),
        'bas_render_step': ([C.POINTER(StepJob), vp], i),
        'bas_render_routed': ([vp, ll, ll, i, ll, i, i, i, vp, vp, vp, i, vp, ll, ll, vp, ll, vp, i, vp, ll, C.POINTER(Route), vp], i),
        'bas_peer_signal': ([vp, i, i, C.c_uint, vp], i),
        'bas_peer_reduce': ([vp, i, ll, ll, vp, i, ll, ll, vp, C.c_uint, vp, i, vp, vp, vp], i),
        'bas_peer_wait': ([vp, i, C.c_uint, vp], i),
        'bas_render_set_trace': ([vp], i),
        'bas_peer_stream_wait': ([vp, i, C.c_uint, vp], i),
        'bas_render_workspace_bytes': ([], ll),
        'bas_normalise': ([vp, ll, vp, vp], i),
        'bas_peak': ([vp, ll, vp, vp], i),
        'bas_copy_2d': ([vp, ll, vp, ll, ll, ll, i, vp], i),
        'bas_host_register': ([vp, ll], i),
        'bas_host_unregister': ([vp], i),
        'bas_pipeline_arena_bytes': ([i, ll, i, i, i, ll, i, C.POINTER(ll)], ll),
        'bas_pipeline_upload': ([C.POINTER(PipelineJob), i, C.POINTER(ll)], i),
        'bas_pipeline_phase': ([C.POINTER(PipelineJob), i, i, ll, ll, ll, ll], i),
        'bas_pipeline_trace': ([i, C.c_char_p, C.c_size_t], i),
        'bas_memset': ([vp, i, ll, vp], i),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of step
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.bas_abi_version() != ABI_VERSION:
        raise ImportError('libbas_b200.so has ABI %d, expected %d: rebuild it' % (lib.bas_abi_version(), ABI_VERSION))
    return lib, tuple(sigs)


lib, EXPORTS = _load()


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib.bas_last_error(buf, 512)
    return buf.value.decode(errors='replace')


def decode_status(words):
    """(error bits of the earliest failing point, its index) from the two status words of a plan
    launch (include/bas_b200.h): the reference raises at the first bad trajectory point."""
    if not int(words[0]):
        return 0, 0
    packed = ~int(words[1]) & 0xffffffff             # stored complemented, so that zeroed words mean "no error"
    return packed & 7, packed >> 3


def check(rc: int, what: str) -> None:
    """Raise on a non-zero status from a launch-type entry point."""
    if rc != 0:
        raise BasError('%s failed (status %d): %s' % (what, rc, last_error()))


_workspaces = {}


def render_workspace(torch, device):
    """Scratch for bas_render's tile splitting (~19 MB), one per (device, host thread, current stream):
    concurrent renders must not share it (include/bas_b200.h)."""
    import threading
    key = (device.index, threading.get_ident(), torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.empty(int(lib.bas_render_workspace_bytes()), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def require_device():
    """The product path has no CPU implementation: fail loudly without a CUDA device."""
    import torch
    if not torch.cuda.is_available() or lib.bas_device_count() < 1:
        raise BasError('binaural-audio-synthesis_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch
