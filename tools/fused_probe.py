"""Render-only time of the tiled kernel per tile shape, with the filter rows synthesised in-kernel (fused) and
copied from HBM (two-kernel path; bas_ir_synth timed separately).   python tools/fused_probe.py [n_src] [K] [U] [mix]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 16
keep = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ups = int(sys.argv[3]) if len(sys.argv) > 3 else 8
mix = bool(int(sys.argv[4])) if len(sys.argv) > 4 else n_src > 1
ah, cabi, lib = bas.apply_hrtf, bas._cabi, bas._cabi.lib
bank = bench.make_bank(ups, keep)
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
shapes = [(0, 0, 0)] + list(cabi.TILED_SHAPES)
FUSED_SHAPES = {(0, 0, 0), (4, 2, 2), (8, 2, 1)}
for fused in (True, False):
    ah.FUSED = fused
    for tw, ns, ctas in shapes:
        if fused and (tw, ns, ctas) not in FUSED_SHAPES:
            continue
        variant = cabi.render_variant(tw, ns, ctas, 0) if tw else 0
        try:
            job = ah.DeviceRender(torch, bdev, x, n_in, 512, 32, elev, azim, cabi.AZ_F64, mix, variant)
            job.plan(st)
            ms = timed(lambda: job.render(st, 0, n_out, out.data_ptr(), stride))
            res['%s tw%d ns%d ctas%d' % ('fused' if fused else 'plain', tw, ns, ctas)] = round(1e3 * ms / n_src, 1)
        except Exception as e:
            res['%s tw%d ns%d ctas%d' % ('fused' if fused else 'plain', tw, ns, ctas)] = str(e)[:60]
        del job
ah.FUSED = False
job = ah.DeviceRender(torch, bdev, x, n_in, 512, 32, elev, azim, cabi.AZ_F64, mix, 0)
res['plan+ir_synth (us per source)'] = round(1e3 * timed(lambda: job.plan(st)) / n_src, 1)
ah.FUSED = True
job = ah.DeviceRender(torch, bdev, x, n_in, 512, 32, elev, azim, cabi.AZ_F64, mix, 0)
res['plan only (us per source)'] = round(1e3 * timed(lambda: job.plan(st)) / n_src, 1)
print(json.dumps({'n_src': n_src, 'K': k, 'U': ups, 'mix': mix, 'unit': 'us per 60 s source, render only', 'results': res}, indent=1))
