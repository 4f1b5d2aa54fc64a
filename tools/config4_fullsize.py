"""BASELINE config 4 at full size on one GPU: one 1-hour 48 kHz source (172.8 M samples, 337,501 chunk
boundaries), device resident; windows against the oracle, and the 8-way time cut against the whole."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
from oracle import binaural_oracle as oracle
import bench

bas.apply_hrtf.PROGRESS = False
bank = bench.make_bank(bas)
fs, n = 48000, 3600 * 48000
g = torch.Generator(device='cuda').manual_seed(4)
x = (0.05 * torch.randn(n, device='cuda', generator=g)).reshape(1, n).contiguous()
k_ = 2 * np.pi / (60 * fs)

def spiral(t):
    t = np.asarray(t, dtype=np.float64)
    return (np.deg2rad(-45.0) + (np.deg2rad(135.0) / n) * t, (15 * 60 * k_ * t / 60) % (2 * np.pi))
spiral.vectorized = True
taps, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
torch.cuda.synchronize(); t0 = time.perf_counter()
y = bas.render_sources(x, 512, 32, [spiral], bank, normalise=False, return_device=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print('1 h source: %d output pairs in %.1f ms (first call, includes trajectory evaluation on the host)' % (n_out, (t1 - t0) * 1e3))
assert y.shape == (1, 2, n_out) and bool(torch.isfinite(y).all())
xh = x[0].cpu().numpy()
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
for p0 in (0, 512 * 100_000 + 7, n_out - 700):
    n0 = max(0, (p0 - taps) // 512 * 512)
    seg = xh[n0:min(n, p0 + 600)]
    want = oracle.make_signal_move_2d(seg, 512, 32, lambda t: spiral(np.float64(t + n0)), bank).T[:, p0 - n0:p0 - n0 + 600]
    got = y[0, :, p0:p0 + want.shape[1]].cpu().numpy()
    print('window at %d: rel-L2 %.2e' % (p0, rel(got.astype(np.float64), want)))
    assert rel(got.astype(np.float64), want) <= 1e-5
dist = bas.distributed
shape = bas._cabi.render_variant(4, 1, 3, 1, split=False)
whole = bas.render_sources(x, 512, 32, [spiral], bank, normalise=False, return_device=True, variant=shape)[0]
for r, (p0, p1) in enumerate(dist.time_segments(n_in, 512, taps, 8)):
    n0, n1 = dist.segment_inputs(p0, p1, n_in, 512, taps)
    window = torch.zeros((1, n1 - n0), dtype=torch.float32, device='cuda')
    window[0, :min(n1, n) - n0] = x[0, n0:min(n1, n)]
    seg = bas.render_sources(window, 512, 32, [dist._shift_trajectory(spiral, n0)], bank, normalise=False, return_device=True,
                             time_range=(p0 - n0, min(p1, n1 + taps - 1) - n0), variant=shape)[0]
    same = bool(torch.equal(seg, whole[:, p0:p0 + seg.shape[1]]))
    print('segment %d [%d, %d): equals the one-piece render: %s' % (r, p0, p1, same))
    assert same
print('ok')
