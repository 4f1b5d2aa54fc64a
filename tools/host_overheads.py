"""Host-side cost of the building blocks of the e2e call (microbenchmarks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
from binaural_audio_synthesis_b200 import _cabi
lib = _cabi.lib

def t(name, fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    print('%-44s %8.2f us' % (name, dt * 1e6))

N = 2 * 2646271
keep = []
def pin_big():
    keep.append(torch.empty(N, dtype=torch.float32, pin_memory=True))
    if len(keep) > 2: keep.pop(0)
t('torch.empty pinned 21MB (2 alive)', pin_big, 50)
t('torch.empty pinned 1KB', lambda: torch.empty(256, dtype=torch.float32, pin_memory=True))
t('torch.empty device 21MB', lambda: torch.empty(N, dtype=torch.float32, device='cuda'))
t('torch.zeros device 1 float', lambda: torch.zeros(1, dtype=torch.float32, device='cuda'))
t('np.empty 21MB', lambda: np.empty(N, dtype=np.float32))
s2 = torch.cuda.Stream()
main = torch.cuda.current_stream()
def evt():
    e = torch.cuda.Event(); e.record(main); s2.wait_event(e)
t('Event create+record+wait', evt)
t('torch.cuda.current_stream()', lambda: torch.cuda.current_stream())
d = torch.empty(1024, device='cuda'); h = torch.empty(1024, pin_memory=True)
t('bas_copy_2d 4KB d2h (ctypes)', lambda: lib.bas_copy_2d(h.data_ptr(), 4096, d.data_ptr(), 4096, 4096, 1, 0, main.cuda_stream))
t('tensor.copy_ 4KB d2h non_blocking', lambda: h.copy_(d, non_blocking=True))
t('data_ptr()', lambda: d.data_ptr())
t('stream.synchronize (idle)', lambda: main.synchronize())
import bench
bas.apply_hrtf.PROGRESS = False
bank = bench.make_bank(bas)
n = 60 * 44100
x = torch.from_numpy(bench.pink_noise(n, 2)).pin_memory().numpy()
traj = bench.lissajous(0)
times = np.arange(0, n + 513, 512, dtype=np.int64)
t('trajectory fn (5169 pts)', lambda: traj(times))
t('np.arange times', lambda: np.arange(0, n + 513, 512, dtype=np.int64))
for seg in (1 << 20, 2 << 20, 3 << 20, 6 << 20, 64 << 20):
    bas.apply_hrtf.PIPELINE_SEGMENT_BYTES = seg
    t('make_signal_move_2d seg=%dMB' % (seg >> 20), lambda: bas.make_signal_move_2d(x, 512, 32, traj, bank), 30)
