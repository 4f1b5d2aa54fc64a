"""e2e time of make_signal_move_2d (pinned host in, host out) over pipeline settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench
ah = bas.apply_hrtf
ah.PROGRESS = False
bank = bench.make_bank(bas)
n = 60 * 44100
x = torch.from_numpy(bench.pink_noise(n, 2)).pin_memory().numpy()
traj = bench.lissajous(0)
def run(reps=200):
    for _ in range(10):
        y = bas.make_signal_move_2d(x, 512, 32, traj, bank)
    t0 = time.perf_counter()
    for _ in range(reps):
        y = bas.make_signal_move_2d(x, 512, 32, traj, bank)
    return (time.perf_counter() - t0) / reps * 1e6
for phases in ((1.0,), (0.3, 1.0), (0.5, 1.0), (0.2, 1.0), (0.15, 0.5, 1.0), (0.1, 0.3, 0.6, 1.0)):
    for seg in (4, 8, 16):
        ah.PIPELINE_PHASES = phases
        ah.PIPELINE_SEGMENT_BYTES = seg << 20
        print('phases %-22s segment %2d MB: %7.1f us' % (phases, seg, run()), flush=True)
