"""FP32 pipe ceilings on this GPU: dependent-chain probes and FIR-shaped register streams at
different numbers of resident warps per SM.   python tools/probe_fma.py"""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from binaural_audio_synthesis_b200 import _cabi

sys.path.insert(0, os.path.join(ROOT, 'tools'))
import probe_lib
lib = probe_lib.load()
dev = torch.device('cuda', 0)
stream = torch.cuda.current_stream().cuda_stream
sink = torch.empty(148 * 64 * 256, dtype=torch.float32, device=dev)


def run(mode, blocks, threads, iters, fma_per_thread_iter):
    for _ in range(2):
        _cabi.check(lib.bas_probe_fma(mode, blocks, threads, iters, sink.data_ptr(), stream), 'probe')
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(lib.bas_probe_fma(mode, blocks, threads, iters, sink.data_ptr(), stream), 'probe')
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return blocks * threads * iters * fma_per_thread_iter / (best * 1e-3) / 1e12


out = {}
for name, mode, per in (('chain_f32', 0, 32), ('chain_f32x2', 1, 32)):
    out[name] = {w: round(run(mode, 148 * (w // 4), 128, 4096, per), 2) for w in (4, 8, 16, 32)}
for name, mode in (('fir_f32x2', 2), ('fir_f32', 3), ('diag_f32x2', 4), ('diag_f32', 5)):
    out[name] = {w: round(run(mode, 148 * (w // 4), 128, 64, 2048), 2) for w in (4, 8, 12, 16, 20)}
print(json.dumps(out, indent=1))
