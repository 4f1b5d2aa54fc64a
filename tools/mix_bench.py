"""BASELINE config 3 / 5 shape on one GPU: n_src device-resident 60 s sources mixed to one binaural
output (the per-rank share of the by-source sharding), with the filter rows synthesised inside the render
kernel (fused, the default) and by a separate bas_ir_synth launch.   python tools/mix_bench.py [n_src] [K] [U]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
import bench

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 64
keep = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ups = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ah = bas.apply_hrtf
bank = bench.make_bank(ups, keep)
fs = 44100
n = 60 * fs
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
x = (0.05 / 8 * torch.randn((n_src, n_in), device='cuda')).contiguous()
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
dirs = [bench.lissajous(1 + s, fs)(times) for s in range(n_src)]
elev = torch.from_numpy(np.stack([np.broadcast_to(d[0], times.shape) for d in dirs])).cuda()
azim = torch.from_numpy(np.stack([d[1] for d in dirs])).cuda()
pre = (elev, azim, bas._cabi.AZ_F64)
variant = int(os.environ.get('BAS_VARIANT', '0'), 0)
out = {}
for label, fused in (('fused filter synthesis', True), ('separate ir_synth launch', False)):
    ah.FUSED = fused
    for _ in range(2):
        y = ah.render_sources(x, 512, 32, pre, bank, mix=True, return_device=True, variant=variant)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        y = ah.render_sources(x, 512, 32, pre, bank, mix=True, return_device=True, variant=variant)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    out[label] = {'ms': round(ms, 3), 'source_samples_per_s': round(n_src * n_out / ms * 1e3 / 1e9, 2), 'us_per_source': round(ms * 1e3 / n_src, 1)}
print(json.dumps({'n_src': n_src, 'K': k, 'U': ups, 'unit': 'G source-samples/s', 'results': out}, indent=1))
