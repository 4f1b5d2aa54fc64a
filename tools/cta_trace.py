"""Per-CTA time stamps of the tiled render kernel (bas_render_set_trace): how evenly do the CTAs of one launch
finish?  One GPU.   python tools/cta_trace.py [n_src ...]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench

ah, cabi, lib = bas.apply_hrtf, bas._cabi, bas._cabi.lib
dev = torch.device('cuda', 0)
bank = bench.make_bank(16, 256)
bdev = ah._device_bank(bank)
n = 60 * 44100
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
stride = (n_out + 3) // 4 * 4
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
st = torch.cuda.current_stream().cuda_stream
trace = torch.zeros(4 * 148 * 4 * 4, dtype=torch.int64, device=dev)
variant = int(os.environ.get('BAS_TRACE_VARIANT', '0'), 0)          # a bas_render variant word (cabi.render_variant), 0: the library's choice
for n_src in [int(a) for a in sys.argv[1:]] or [64, 8, 1]:
    mix = n_src > 1
    x = torch.randn((n_src, n_in), device=dev) * 0.01
    dirs = [bench.lissajous(s)(times) for s in range(n_src)]
    elev = torch.from_numpy(np.stack([d[0] for d in dirs])).to(dev)
    azim = torch.from_numpy(np.stack([d[1] for d in dirs])).to(dev)
    job = ah.DeviceRender(torch, bdev, x, n_in, 512, 32, elev.reshape(-1), azim.reshape(-1), cabi.AZ_F64, mix, variant)
    out = torch.zeros((2, stride) if mix else (1, 2, stride), dtype=torch.float32, device=dev)
    job.plan(st)
    for rep in range(3):
        job.render(st, 0, n_out, out.data_ptr(), stride)
    torch.cuda.synchronize()
    rows = []
    for rep in range(3):
        trace.zero_()
        lib.bas_render_set_trace(trace.data_ptr())
        job.render(st, 0, n_out, out.data_ptr(), stride)
        torch.cuda.synchronize()
        lib.bas_render_set_trace(None)
        t = trace.cpu().numpy().reshape(-1, 4)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        start, end, smid, items = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2], t[:, 3]
        dur = end - start
        rows.append({'ctas': len(t), 'kernel_us': round(float(end.max()), 1), 'start_us_max': round(float(start.max()), 1),
                     'end_us_min_p10_median_p90_max': [round(float(v), 1) for v in np.percentile(end, [0, 10, 50, 90, 100])],
                     'duration_us_min_median_max': [round(float(v), 1) for v in np.percentile(dur, [0, 50, 100])],
                     'idle_tail_mean_us': round(float((end.max() - end).mean()), 1), 'items_min_max': [int(items.min()), int(items.max())]})
    order = np.argsort(end)
    print(json.dumps({'n_src': n_src, 'variant': hex(variant), 'launches': rows,
                      'earliest_10': [[int(smid[i]), round(float(end[i]), 1), int(items[i])] for i in order[:10]],
                      'latest_10': [[int(smid[i]), round(float(end[i]), 1), int(items[i])] for i in order[-10:]],
                      'end_by_sm_mod_2': [round(float(end[smid % 2 == p].mean()), 1) for p in (0, 1)],
                      'end_by_cta_quarter': [round(float(end[(np.arange(len(end)) * 4 // len(end)) == q].mean()), 1) for q in range(4)]}))
    del job, x, out
