"""Time every compiled tile shape of the render kernel on the bench workload (device-resident).
    python tools/tune_render.py [n_src] [K] [U]
"""
import os
import sys
import json

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import binaural_audio_synthesis_b200 as bas
from binaural_audio_synthesis_b200 import _cabi
import bench

lib = _cabi.lib


def main():
    n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    keep = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ups = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    mix = int(sys.argv[4]) if len(sys.argv) > 4 else (1 if n_src > 1 else 0)
    only_tw = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    only_names = sys.argv[6].split(',') if len(sys.argv) > 6 else None
    dev = torch.device('cuda', 0)
    f = bas.bank_synth.build_bank(ups, seed=0)

    class bank:
        upsampling = ups
        diffs_left, diffs_right = f['diffs_left'], f['diffs_right']
        irs_left, irs_right = f['irs_left'][:, :keep * ups], f['irs_right'][:, :keep * ups]
    bdev = bas.apply_hrtf._device_bank(bank)
    seconds = float(os.environ.get('BAS_SECONDS', '60'))
    n = int(seconds * 44100)
    k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
    n_pts = n_in // 512 + 1
    stream = torch.cuda.current_stream().cuda_stream
    x = (0.05 * torch.randn((n_src, n_in), device=dev)).contiguous()
    pitch = lib.bas_filter_row_pitch(k)
    filt = (0.05 * torch.randn((n_src, n_pts, pitch, 2), device=dev)).contiguous()
    out = torch.empty((1 if mix else n_src, 2, (n_out + 4) // 4 * 4), device=dev)
    peaks = torch.zeros(n_src, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    results = {}
    workspace = _cabi.render_workspace(torch, dev)
    for (tw, ns, ctas) in ((0, 0, 0),) + tuple(_cabi.TILED_SHAPES):
        for parts in (0, 1, 2, 4):
            for split in (False, True):
                if (tw == 0) != (parts == 0) or (tw and tw % parts):
                    continue
                if only_tw and tw != only_tw:
                    continue
                if only_names and '%dx%dx%d/p%d%s' % (tw, ns, ctas, parts, '/split' if split else '') not in only_names:
                    continue
                variant = _cabi.render_variant(tw, ns, ctas, parts, split)
                name = '%dx%dx%d/p%d%s' % (tw, ns, ctas, parts, '/split' if split else '')

                def run():
                    return lib.bas_render(x.data_ptr(), n_in, n_in, n_src, n_in, 512, 32, k, filt.data_ptr(), None, 0, n_out,
                                          out.data_ptr(), out.shape[-1], mix, peaks.data_ptr(), variant, workspace.data_ptr(), workspace.numel(), stream)
                rc = run()
                if rc != 0:
                    results[name] = 'rc=%d %s' % (rc, _cabi.last_error())
                    continue
                torch.cuda.synchronize()
                times = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run()
                    e1.record()
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
                ms = float(np.median(times))
                results[name] = {'ms': round(ms, 4), 'tfma_s': round(2.0 * k * n_in * n_src / ms / 1e9, 2), 'Gpairs_s': round(n_out * n_src / ms / 1e6, 2)}
    print(json.dumps({'seconds': seconds, 'mix': mix, 'n_src': n_src, 'K': k, 'U': ups, 'results': results}, indent=1))


if __name__ == '__main__':
    main()
