"""Device time of bas_render over sub-ranges of the output, alone and beside a device->host copy."""
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
side = torch.cuda.Stream()
host = torch.empty((2, stride), pin_memory=True)

def render(pa, pb):
    _cabi.check(lib.bas_render(x.data_ptr(), n_in, n_in, 1, n_in, 512, 32, k, filt.data_ptr(), None, pa, pb - pa,
                               out.data_ptr() + 4 * pa, stride, 0, peaks.data_ptr(), 0, ws.data_ptr(), ws.numel(), main.cuda_stream), 'render')

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main); fn(); e1.record(main); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

print('whole             %.1f us' % timed(lambda: render(0, n_out)))
for parts in (2, 3, 5, 8):
    step = (n_out // parts + 8191) // 8192 * 8192
    cuts = list(range(0, n_out, step)) + [n_out]
    each = [timed(lambda a=a, b=b: render(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
    allof = timed(lambda: [render(a, b) for a, b in zip(cuts[:-1], cuts[1:])])
    def with_copy():
        with torch.cuda.stream(side):
            host.copy_(out[0], non_blocking=True)
        for a, b in zip(cuts[:-1], cuts[1:]):
            render(a, b)
    print('%d parts: each %s   back-to-back %.1f us   beside a 21 MB D2H %.1f us' % (parts, ['%.1f' % t for t in each], allof, timed(with_copy)))
    torch.cuda.synchronize()
