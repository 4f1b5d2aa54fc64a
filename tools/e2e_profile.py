"""Where the end-to-end call spends its host time (cProfile of make_signal_move_2d, pinned input)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench

bas.apply_hrtf.PROGRESS = False
bank = bench.make_bank(bas)
n = 60 * 44100
x = torch.from_numpy(bench.pink_noise(n, 2)).pin_memory().numpy()
traj = bench.lissajous(0)
for _ in range(3):
    bas.make_signal_move_2d(x, 512, 32, traj, bank)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    y = bas.make_signal_move_2d(x, 512, 32, traj, bank)
print('ms per call', (time.perf_counter() - t0) / 20 * 1e3)
# raw copies for reference
xd = torch.empty(n, device='cuda')
xt = torch.from_numpy(x)
out_d = torch.empty((2, n + 300), device='cuda')
out_h = torch.empty((2, n + 300), pin_memory=True)
for name, fn in (('h2d pinned 10.6MB', lambda: xd.copy_(xt, non_blocking=True)), ('d2h pinned 21MB', lambda: out_h.copy_(out_d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / 20 * 1e3, 'ms')
t0 = time.perf_counter()
for _ in range(20):
    h = torch.empty((2, n + 300), pin_memory=True)
print('pinned alloc', (time.perf_counter() - t0) / 20 * 1e3, 'ms')
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    y = bas.make_signal_move_2d(x, 512, 32, traj, bank)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
