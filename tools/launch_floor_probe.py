"""Device time of bas_render for growing output ranges (how much of a short render is fixed cost)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
from binaural_audio_synthesis_b200 import _cabi
import bench
lib = _cabi.lib
dev = torch.device('cuda', 0)
bank = bench.make_bank(bas)
n = 60 * 44100
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
n_pts = n_in // 512 + 1
pitch = lib.bas_filter_row_pitch(k)
x = (0.05 * torch.randn((1, n_in), device=dev)).contiguous()
filt = (0.05 * torch.randn((1, n_pts, pitch, 2), device=dev)).contiguous()
stride = (n_out + 3) // 4 * 4
out = torch.empty((1, 2, stride), device=dev)
peaks = torch.zeros(1, device=dev)
ws = _cabi.render_workspace(torch, dev)
main = torch.cuda.current_stream()

def render(count, variant):
    _cabi.check(lib.bas_render(x.data_ptr(), n_in, n_in, 1, n_in, 512, 32, k, filt.data_ptr(), None, 0, count,
                               out.data_ptr(), stride, 0, peaks.data_ptr(), variant, ws.data_ptr(), ws.numel(), main.cuda_stream), 'render')

def timed(fn, reps=9):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main); fn(); fn(); fn(); fn(); e1.record(main); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 4)
    return float(np.median(ts))

for name, variant in (('4x1x3', _cabi.render_variant(4, 1, 3, 1, False)), ('4x2x2', _cabi.render_variant(4, 2, 2, 1, False)), ('8x2x1', _cabi.render_variant(8, 2, 1, 1, False))):
    tile = (8 if name.startswith('8') else 4) * 1024
    per_wave = 148 * (3 if name == '4x1x3' else 2 if name == '4x2x2' else 1)
    for tiles in (1, 37, 148, per_wave, per_wave + 1, 2 * per_wave, 3 * per_wave, 4 * per_wave):
        count = min(n_out, tiles * tile)
        render(count, variant); torch.cuda.synchronize()
        print('%s %5d tiles (%.2f waves): 4 back-to-back launches, warm L2: %6.1f us each' % (name, tiles, tiles / per_wave, timed(lambda: render(count, variant))), flush=True)
