"""Per-item versus per-block cost inside the render kernel: one tile (a lone CTA), and one full wave, at
K = 256, 512 and 1024 taps (8, 16, 32 blocks per item): time = fixed + blocks * per_block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from binaural_audio_synthesis_b200 import _cabi
lib = _cabi.lib
dev = torch.device('cuda', 0)
n_in = 60 * 44100 // 512 * 512 + 512
n_pts = n_in // 512 + 1
x = (0.05 * torch.randn((1, n_in), device=dev)).contiguous()
ws = torch.empty(int(lib.bas_render_workspace_bytes()), dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream()
peaks = torch.zeros(1, device=dev)

def timed(fn, reps=9):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main); fn(); fn(); fn(); fn(); e1.record(main); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 4)
    return float(np.median(ts))

rows = {}
for k in (256, 512, 1024):
    pitch = lib.bas_filter_row_pitch(k)
    filt = (0.05 * torch.randn((1, n_pts, pitch, 2), device=dev)).contiguous()
    n_out = n_in + k - 1
    stride = (n_out + 3) // 4 * 4
    out = torch.empty((1, 2, stride), device=dev)
    for name, variant, tile, per_wave in (('4x1x2', _cabi.render_variant(4, 1, 2, 1, False), 4096, 296), ('4x1x3', _cabi.render_variant(4, 1, 3, 1, False), 4096, 444)):
        for tiles in (1, per_wave):
            count = tiles * tile
            def run():
                return lib.bas_render(x.data_ptr(), n_in, n_in, 1, n_in, 512, 32, k, filt.data_ptr(), None, 0, count, out.data_ptr(), stride, 0,
                                      peaks.data_ptr(), variant, ws.data_ptr(), ws.numel(), main.cuda_stream)
            rc = run()
            if rc != 0:
                print(name, k, tiles, 'rc', rc, _cabi.last_error()); continue
            torch.cuda.synchronize()
            rows[(name, tiles == 1, k)] = timed(run)
for name in ('4x1x2', '4x1x3'):
    for lone in (True, False):
        t = [rows.get((name, lone, k)) for k in (256, 512, 1024)]
        if None in t:
            print(name, 'lone tile' if lone else 'full wave', t); continue
        per_block_a = (t[1] - t[0]) / 8
        per_block_b = (t[2] - t[1]) / 16
        print('%s %-9s K=256 %.1f us, K=512 %.1f us, K=1024 %.1f us -> per block %.2f / %.2f us (%d / %d cycles), fixed %.1f us' % (
            name, 'lone tile' if lone else 'full wave', t[0], t[1], t[2], per_block_a, per_block_b, per_block_a * 1965, per_block_b * 1965, t[0] - 8 * per_block_a))
