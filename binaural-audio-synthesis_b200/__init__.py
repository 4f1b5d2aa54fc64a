"""B200-native moving-source binaural renderer: drop-in for the hot path of
mbjd/binaural-audio-synthesis' apply_hrtf.py (load_irs_and_delaydiffs, interpolate_2d,
delay_compensated_interpolation_with_delaydiff, make_signal_move_2d) on hand-written sm_100a CUDA
kernels behind the C ABI of include/bas_b200.h.  No CPU fallback.

Import as `binaural_audio_synthesis_b200` (see the alias package of that name).
"""
from . import _cabi, sphere, apply_hrtf, bank_synth, bank_builder, distributed          # noqa: F401
from ._cabi import BasError                                                # noqa: F401
from .apply_hrtf import (                                                  # noqa: F401
    load_irs_and_delaydiffs,
    delay_compensated_interpolation_with_delaydiff,
    delay_compensated_interpolation,
    delay_compensated_interpolation_easy,
    delay_signal_float,
    interpolate_2d,
    interpolate_2d_deg,
    interpolate_2d_batch,
    make_signal_move_2d,
    make_signal_move,
    render_sources,
    render_geometry,
    suggest_chunk_sizes,
    evaluate_trajectory,
    plan_points_host,
)

__all__ = [
    'load_irs_and_delaydiffs', 'delay_compensated_interpolation_with_delaydiff',
    'delay_compensated_interpolation', 'delay_compensated_interpolation_easy', 'delay_signal_float', 'interpolate_2d',
    'interpolate_2d_deg', 'interpolate_2d_batch', 'make_signal_move_2d', 'make_signal_move', 'render_sources',
    'render_geometry', 'suggest_chunk_sizes', 'evaluate_trajectory', 'plan_points_host', 'sphere', 'bank_synth', 'bank_builder', 'distributed',
    'BasError',
]
