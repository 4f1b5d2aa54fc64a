"""Render time of the tiled kernel at subchunksize 16 / 32 / 64 (within 1.2 x of the S = 32 time).
    python tools/subchunk_probe.py [n_src]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ah, cabi = bas.apply_hrtf, bas._cabi
bank = bench.make_bank(8, 256)
bdev = ah._device_bank(bank)
fs, n = 44100, 60 * 44100
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
dev = torch.device('cuda', 0)
x = (0.05 / 8 * torch.randn((n_src, n_in), device=dev)).contiguous()
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
dirs = [bench.lissajous(1 + s, fs)(times) for s in range(n_src)]
elev = torch.from_numpy(np.stack([d[0] for d in dirs])).to(dev).reshape(-1)
azim = torch.from_numpy(np.stack([d[1] for d in dirs])).to(dev).reshape(-1)
stride = (n_out + 3) // 4 * 4
mix = n_src > 1
out = torch.empty((1 if mix else n_src, 2, stride), device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


res = {}
for fused in (True, False):
    ah.FUSED = fused
    for sub in (16, 32, 64, 8):
        variant = cabi.RENDER_AUTO
        job = ah.DeviceRender(torch, bdev, x, n_in, 512, sub, elev, azim, cabi.AZ_F64, mix, variant)
        job.plan(st)
        ms = timed(lambda: job.render(st, 0, n_out, out.data_ptr(), stride), reps=8 if sub != 8 else 2)
        res['%s S=%d%s' % ('fused' if fused else 'two-kernel', sub, ' (generic kernel)' if sub == 8 else '')] = round(1e3 * ms / n_src, 1)
print(json.dumps({'n_src': n_src, 'unit': 'us per 60 s source, render only (K = 256, chunksize 512)', 'results': res}, indent=1))
