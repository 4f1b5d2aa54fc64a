"""Device and host timeline of one make_signal_move_2d call (host array in, host array out)."""
import os, sys, time
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
if len(sys.argv) > 1:
    bas.apply_hrtf.PIPELINE_SEGMENT_BYTES = int(float(sys.argv[1]) * (1 << 20))
for _ in range(5):
    bas.make_signal_move_2d(x, 512, 32, traj, bank)
torch.cuda.synchronize()
import ctypes
lib = bas._cabi.lib
lib.bas_pipeline_trace(1, None, 0)
t0 = time.perf_counter()
y = bas.make_signal_move_2d(x, 512, 32, traj, bank)
t1 = time.perf_counter()
buf = ctypes.create_string_buffer(1 << 16)
lib.bas_pipeline_trace(0, buf, len(buf))
print('call %.1f us' % ((t1 - t0) * 1e6))
print(buf.value.decode())
with bench.ClockSampler(0) as clocks:
    t0 = time.perf_counter()
    for _ in range(300):
        y = bas.make_signal_move_2d(x, 512, 32, traj, bank)
    dt = (time.perf_counter() - t0) / 300
print('loop of 300 calls: %.1f us per call; clocks %s' % (dt * 1e6, clocks.summary()))
