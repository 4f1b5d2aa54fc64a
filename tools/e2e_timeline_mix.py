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

# the same time ranges rendered from device-resident input, back to back (GPU busy, no copies in flight): is a phase's
# render slower inside the pipeline than on its own?
import subprocess
xd = torch.from_numpy(x).cuda()
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
cuts = [0, 212992, 532480, 933888, 1327104, 1720320, 2121728, 2441216, n_out]
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
dirs = [bench.lissajous(s)(times) for s in range(n_src)]
pre = (np.stack([d_[0] for d_ in dirs]), np.stack([d_[1] for d_ in dirs]), bas._cabi.AZ_F64)
pre_d = (torch.from_numpy(pre[0]).cuda(), torch.from_numpy(pre[1]).cuda(), pre[2])
for rep in range(2):
    row = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if rep == 1:
            time.sleep(0.02)                                 # let the GPU go idle first, like between the phases of a call
        e0.record()
        bas.render_sources(xd, 512, 32, pre_d, bank, mix=True, normalise=False, return_device=True, time_range=(a, b))
        e1.record(); torch.cuda.synchronize()
        row.append(round(e0.elapsed_time(e1) * 1e3))
    print('device-resident ranges (plan of all points + render), us, %s: %s' % ('after 20 ms idle' if rep else 'back to back', row))
q = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,pstate,power.draw', '--format=csv,noheader'], capture_output=True, text=True)
print('idle now:', q.stdout.strip())
