"""Do ir_synth (L2 bound) and render (FP32-pipe bound) of independent sources overlap on one GPU?
Times N synth launches and N render launches back to back on one stream, and on two streams."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import binaural_audio_synthesis_b200 as bas
from binaural_audio_synthesis_b200 import _cabi
import bench
lib = _cabi.lib
dev = torch.device('cuda', 0)
bank = bench.make_bank(bas)
bdev = bas.apply_hrtf._device_bank(bank)
n = 60 * 44100
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
n_pts = n_in // 512 + 1
pitch = lib.bas_filter_row_pitch(k)
n_src = 16
x = (0.05 * torch.randn((n_src, n_in), device=dev)).contiguous()
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
e, a = bench.lissajous(1)(times)
elev = torch.from_numpy(np.tile(np.broadcast_to(e, times.shape), n_src)).to(dev)
azim = torch.from_numpy(np.tile(a, n_src)).to(dev)
terms = torch.empty(n_src * n_pts * 256, dtype=torch.uint8, device=dev)
status = torch.zeros(2, dtype=torch.int32, device=dev)
filt = [torch.empty((n_src * n_pts, pitch, 2), dtype=torch.float32, device=dev) for _ in range(2)]
stride = (n_out + 3) // 4 * 4
out = torch.empty((2, stride), device=dev)
peaks = torch.zeros(n_src, device=dev)
s0, s1 = (torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)) if os.environ.get('BAS_PRIO') else (torch.cuda.current_stream(), torch.cuda.Stream())
torch.cuda.set_stream(s0)
ws = [torch.empty(int(lib.bas_render_workspace_bytes()), dtype=torch.uint8, device=dev) for _ in range(2)]
lib.bas_plan_build(bdev.diffs[0].data_ptr(), bdev.diffs[1].data_ptr(), 8, k * 8, elev.data_ptr(), azim.data_ptr(), None, 1, n_src * n_pts,
                   terms.data_ptr(), None, status.data_ptr(), s0.cuda_stream)
variant = int(os.environ.get('BAS_VARIANT', '0'), 0)

def synth(st, b):
    _cabi.check(lib.bas_ir_synth(bdev.bank_pp.data_ptr(), 8, k, terms.data_ptr(), n_src * n_pts, _cabi.IR_ROWS, filt[b].data_ptr(), k, st.cuda_stream), 'synth')

def render(st, b):
    _cabi.check(lib.bas_render(x.data_ptr(), n_in, n_in, n_src, n_in, 512, 32, k, filt[b].data_ptr(), None, 0, n_out, out.data_ptr(), stride, 1,
                               peaks.data_ptr(), variant, ws[b].data_ptr(), ws[b].numel(), st.cuda_stream), 'render')

def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s0); s1.wait_event(e0)
    fn()
    ev = torch.cuda.Event(); ev.record(s1); s0.wait_event(ev)
    e1.record(s0); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

synth(s0, 0); synth(s0, 1); torch.cuda.synchronize()
N = 4
t_s = timed(lambda: [synth(s0, 1) for _ in range(N)])
t_r = timed(lambda: [render(s0, 0) for _ in range(N)])
t_both = timed(lambda: [(synth(s1, 1), render(s0, 0)) for _ in range(N)])
t_both_r = timed(lambda: [(render(s0, 0), synth(s1, 1)) for _ in range(N)])
t_both_hi = None
print('variant %#x: %d x synth %.3f ms, %d x render %.3f ms, serial sum %.3f ms, two streams %.3f ms (synth issued first) / %.3f ms (render issued first)' % (
    variant, N, t_s, N, t_r, t_s + t_r, t_both, t_both_r))
