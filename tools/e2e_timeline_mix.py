"""Timeline of one render_sources(host (n_src, N), mix=True) call.   python tools/e2e_timeline_mix.py [n_src]"""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bas.apply_hrtf.PROGRESS = False
bank = bench.make_bank(8, 256)
n = 60 * 44100
x = np.stack([bench.noise_host(n, 100 + s, 0.05 / 8) for s in range(n_src)])
x = torch.from_numpy(x).pin_memory().numpy()
trajs = []
for s in range(n_src):
    f = bench.lissajous(s)
    g = (lambda t, f=f: f(t)); g.vectorized = True
    trajs.append(g)
for _ in range(3):
    bas.render_sources(x, 512, 32, trajs, bank, mix=True)
torch.cuda.synchronize()
lib = bas._cabi.lib
lib.bas_pipeline_trace(1, None, 0)
t0 = time.perf_counter()
y = bas.render_sources(x, 512, 32, trajs, bank, mix=True)
t1 = time.perf_counter()
buf = ctypes.create_string_buffer(1 << 18)
lib.bas_pipeline_trace(0, buf, len(buf))
print('call %.1f us' % ((t1 - t0) * 1e6))
print(buf.value.decode())
# plain H2D bandwidth of the same bytes, one copy
d = torch.empty(x.shape, dtype=torch.float32, device='cuda')
xt = torch.from_numpy(x)
for _ in range(2):
    d.copy_(xt, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(xt, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print('plain pinned H2D of %.0f MB: %.2f ms = %.1f GB/s' % (x.nbytes / 1e6, (t1 - t0) * 1e3, x.nbytes / (t1 - t0) / 1e9))

# steady-state per call: pinned + declared vectorised, and pageable + plain lambdas (what bench.py's e2e leg times)
def loop(xs, fns, reps=8):
    for _ in range(2):
        bas.render_sources(xs, 512, 32, fns, bank, mix=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        y = bas.render_sources(xs, 512, 32, fns, bank, mix=True)
    return (time.perf_counter() - t0) / reps * 1e3
print('pinned + vectorised: %.2f ms per call' % loop(x, trajs))
x_page = np.stack([bench.noise_host(n, 100 + s, 0.05 / 8) for s in range(n_src)])
plain = [bench.lissajous(s) for s in range(n_src)]
print('pageable + plain lambdas: %.2f ms per call' % loop(x_page, plain))
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    bas.render_sources(x_page, 512, 32, plain, bank, mix=True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(14); print(s.getvalue()[:3000])
