"""Multi-GPU sharding of the render path (one process per GPU, torch.distributed).

The reference is single-process; SURVEY.md 8(e) identifies the two ways its hot path shards:

  by source   sources are independent calls of make_signal_move_2d (no cross-call state): every rank
              renders and mixes its own sources, then one SUM reduce of the (2, N_out) fp32 mix over
              NVLink.  No other exchange.
  by time     given the trajectory, chunks are independent except for the K-1 sample FIR tail
              (apply_hrtf.py:450-453).  The signal is cut at multiples of `chunksize`; each rank
              renders the OUTPUT samples of its segment from its inputs plus a K-1 sample input halo
              (output-stationary: no tail exchange), and one MAX all-reduce of the peak implements
              the global normalisation (apply_hrtf.py:462-464).

`local_render` is injectable so the partition / halo / collective logic can be exercised on CPU
ranks (gloo) in tests; the default is the CUDA path and there is no CPU fallback in the product.
"""
from __future__ import annotations

import numpy as np


def shard_sources(n_src: int, rank: int, world: int):
    """Round-robin source indices of `rank` (balanced to within one source)."""
    return list(range(rank, n_src, world))


def time_segments(n_in: int, chunksize: int, ir_length: int, world: int):
    """Cut output samples [0, n_in + K - 1) into `world` contiguous ranges at multiples of
    chunksize (the last range also takes the K-1 tail).  Returns [(p0, p1)] per rank; empty ranges
    are possible when there are fewer chunks than ranks."""
    n_chunks = n_in // chunksize
    cuts = [min(n_chunks, (r * n_chunks + world - 1) // world) * chunksize for r in range(world + 1)]
    n_out = n_in + ir_length - 1
    return [(cuts[r], n_out if r == world - 1 else cuts[r + 1]) for r in range(world)]


def segment_inputs(p0: int, p1: int, n_in: int, chunksize: int, ir_length: int):
    """Input window [n0, n1) a rank needs for outputs [p0, p1): its own samples plus K-1 samples of
    halo on the left, widened to chunk boundaries so that chunk/subchunk phase is preserved."""
    if p1 <= p0:
        return (0, 0)
    n0 = max(0, p0 - (ir_length - 1)) // chunksize * chunksize
    n1 = min(n_in, (min(p1, n_in) + chunksize - 1) // chunksize * chunksize)
    return (n0, max(n1, n0 + chunksize))


def _default_render(*args, **kwargs):
    from .apply_hrtf import render_sources
    return render_sources(*args, **kwargs)


def render_mix_by_source(signals, chunksize, subchunksize, elev_azim_functions, bank, group=None,
                         dst=None, local_render=None, normalise=True):
    """Every rank passes ONLY its own sources (see shard_sources).  Returns the global mix
    (2, N_out): on every rank (all_reduce) when dst is None, else only on rank `dst` (reduce; other
    ranks get None).  The collective runs on the tensor the renderer produced (NCCL on CUDA
    tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist
    render = local_render or _default_render
    if len(signals):
        local = render(signals, chunksize, subchunksize, elev_azim_functions, bank, mix=True,
                       normalise=normalise, return_device=True)
    else:
        local = None
    local = _as_tensor(torch, local)
    if local is not None:
        local = local.contiguous()                  # the renderer returns a view of a padded buffer; NCCL wants dense
    # ranks without sources contribute zeros of the right shape
    shape = torch.tensor([0, 0] if local is None else list(local.shape), dtype=torch.int64,
                         device=local.device if local is not None else _collective_device(torch, dist, group))
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    if local is None:
        local = torch.zeros(tuple(int(v) for v in shape), dtype=torch.float32, device=shape.device)
    if dst is None:
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
        return local
    dist.reduce(local, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return local if dist.get_rank(group) == dst else None


def render_by_time(in_signal, chunksize, subchunksize, elev_azim_function, bank, group=None,
                   local_render=None, gather=True):
    """One long source cut across the ranks of `group` by time.  Every rank holds the whole
    `in_signal` (or at least its own window) and the trajectory.  Returns (segment, (p0, p1)) with
    the rank's normalised output samples (2, p1 - p0), or, with gather=True, the complete
    (N_out, 2) float32 result on every rank - identical to make_signal_move_2d on one device."""
    import torch
    import torch.distributed as dist
    from .apply_hrtf import render_geometry
    render = local_render or _default_render
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = in_signal.shape[0]
    k, n_in, n_out = render_geometry(n, chunksize, subchunksize, bank)
    p0, p1 = time_segments(n_in, chunksize, k, world)[rank]
    n0, n1 = segment_inputs(p0, p1, n_in, chunksize, k)
    seg = None
    if p1 > p0:
        window = in_signal[n0:min(n1, n)]
        if window.shape[0] < n1 - n0:                       # zero padding of apply_hrtf.py:405-406
            pad = np.zeros(n1 - n0, dtype=np.float32)
            pad[:window.shape[0]] = np.asarray(window, dtype=np.float32)
            window = pad
        shifted = _shift_trajectory(elev_azim_function, n0)
        # local output sample p - n0 of the window render equals global sample p for p >= p0:
        # every input it depends on (p-K+1 .. p) is inside the window
        seg = render(window[None, :], chunksize, subchunksize, [shifted], bank, mix=False, normalise=False,
                     return_device=True, time_range=(p0 - n0, min(p1, n1 + k - 1) - n0))
        seg = _as_tensor(torch, seg)[0]
    dev = seg.device if seg is not None else _collective_device(torch, dist, group)
    peak = torch.zeros(1, dtype=torch.float32, device=dev)
    if seg is not None and seg.numel():
        seg = seg.contiguous()
        _peak_into(torch, seg, peak)
    dist.all_reduce(peak, op=dist.ReduceOp.MAX, group=group)            # global apply_hrtf.py:462
    if seg is not None and seg.numel():
        _divide_by_peak(torch, seg, peak)                               # :463-464, only when the peak exceeds 1
    if not gather:
        return seg, (p0, p1)
    full = torch.zeros((2, n_out), dtype=torch.float32, device=dev)
    if seg is not None:
        full[:, p0:p1] = seg
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)            # disjoint ranges: a gather
    out = full.cpu().numpy() if full.is_cuda else full.numpy()
    return out.T


def _peak_into(torch, seg, peak):
    """max |seg| into the one-element tensor `peak`: bas_peak on the device; torch only for the CPU
    tensors a test's injected renderer returns."""
    if seg.is_cuda:
        from ._cabi import check, lib
        check(lib.bas_peak(seg.data_ptr(), seg.numel(), peak.data_ptr(), torch.cuda.current_stream().cuda_stream), 'bas_peak')
    else:
        peak.copy_(seg.abs().max().reshape(1))


def _divide_by_peak(torch, seg, peak):
    if seg.is_cuda:
        from ._cabi import check, lib
        check(lib.bas_normalise(seg.data_ptr(), seg.numel(), peak.data_ptr(), torch.cuda.current_stream().cuda_stream), 'bas_normalise')
    elif float(peak) > 1:
        seg /= peak


def _shift_trajectory(fn, offset: int):
    """Trajectory seen from a window that starts at global sample `offset`."""
    if getattr(fn, 'vectorized', False):
        def shifted(t):
            return fn(np.asarray(t) + offset)
        shifted.vectorized = True
        if hasattr(fn, 'az_kind'):
            shifted.az_kind = fn.az_kind
        return shifted
    return lambda t: fn(t + offset)


def _as_tensor(torch, value):
    if value is None or isinstance(value, torch.Tensor):
        return value
    return torch.from_numpy(np.ascontiguousarray(value))


def _collective_device(torch, dist, group):
    return torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
