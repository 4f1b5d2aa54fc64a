"""How fast the render kernel's 32x32 block runs on its own (no tiles, no TMA, no epilogue)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from binaural_audio_synthesis_b200 import _cabi
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import probe_lib
lib = probe_lib.load()
sink = torch.empty(148 * 3 * 128, dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
out = {}
for ctas, per_sm in ((3, 3), (3, 2), (3, 1), (2, 2), (2, 1)):
    blocks, iters = 148 * per_sm, 40
    for _ in range(2):
        _cabi.check(lib.bas_probe_block(ctas, blocks, iters, sink.data_ptr(), st), 'probe')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _cabi.check(lib.bas_probe_block(ctas, blocks, iters, sink.data_ptr(), st), 'probe'); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    useful = blocks * 128 * iters * 6 * 1024 * 2            # FMAs (two per packed op)
    n_blocks_per_sched = per_sm * iters * 6
    out['compiled for %d CTAs/SM, %d resident (%d warps/scheduler)' % (ctas, per_sm, per_sm)] = {
        'ms': round(ms, 3), 'useful_tfma_s': round(useful / ms / 1e9, 2), 'cycles_per_block_per_scheduler': round(ms * 1e-3 * 1.965e9 / n_blocks_per_sched)}
print(json.dumps(out, indent=1))
