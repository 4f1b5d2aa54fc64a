"""SM clock actually held under FFMA / FFMA2 load (device-side clock64 vs globaltimer), and the FMA rate."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from binaural_audio_synthesis_b200 import _cabi
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import probe_lib
lib = probe_lib.load()
dev = torch.device('cuda', 0)
sink = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)
mhz = torch.zeros(1, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
out = {}
for name, packed in (('ffma', 0), ('ffma2', 1)):
    for iters in (512, 4096, 32768):
        for _ in range(3):
            lib.bas_probe_clock(packed, 148 * 8, 256, iters, sink.data_ptr(), mhz.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.bas_probe_clock(packed, 148 * 8, 256, iters, sink.data_ptr(), mhz.data_ptr(), st); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tf = 148 * 8 * 256 * iters * 32 / (ms * 1e-3) / 1e12
        out['%s iters=%d' % (name, iters)] = {'ms': round(ms, 3), 'tfma_s': round(tf, 2), 'sm_mhz': round(float(mhz), 1),
                                               'fma_per_clk_per_sm': round(tf * 1e12 / (float(mhz) * 1e6) / 148, 1)}
print(json.dumps(out, indent=1))
