import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import binaural_audio_synthesis_b200 as bas, bench
ah, cabi = bas.apply_hrtf, bas._cabi
bank = bench.make_bank(8, 256); bdev = ah._device_bank(bank)
fs, n = 44100, 60 * 44100
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
dev = torch.device('cuda', 0)
x = (0.05 * torch.randn((1, n_in), device=dev)).contiguous()
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
e, a = bench.lissajous(1, fs)(times)
elev = torch.from_numpy(np.ascontiguousarray(e)).to(dev); azim = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
stride = (n_out + 3) // 4 * 4
out = torch.empty((1, 2, stride), device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device=dev)
def timed(fn, reps=10):
    for _ in range(2): fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))
res = {}
for fused in (True, False):
    ah.FUSED = fused
    for (tw, ns, ctas) in ((4, 2, 2), (8, 2, 1)) if fused else ((4, 1, 3), (4, 2, 2), (8, 2, 1)):
        for parts in (1, 2):
            for split in (False, True):
                v = cabi.render_variant(tw, ns, ctas, parts, split=split)
                try:
                    job = ah.DeviceRender(torch, bdev, x, n_in, 512, 32, elev, azim, cabi.AZ_F64, False, v)
                    job.plan(st)
                    res['%s tw%d ns%d ctas%d parts%d %s' % ('fused' if fused else 'plain', tw, ns, ctas, parts, 'split' if split else 'whole')] = round(1e3 * timed(lambda: job.render(st, 0, n_out, out.data_ptr(), stride)), 1)
                except Exception as ex:
                    res['%s tw%d ns%d ctas%d parts%d %s' % ('fused' if fused else 'plain', tw, ns, ctas, parts, 'split' if split else 'whole')] = str(ex)[:50]
print(json.dumps(res, indent=1))
