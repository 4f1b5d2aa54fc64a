"""Multi-GPU sharding of the render path (one process per GPU, torch.distributed).

The reference is single-process; SURVEY.md 8(e) identifies the two ways its hot path shards:

  by source   sources are independent calls of make_signal_move_2d (no cross-call state): every rank
              renders and mixes its own sources, and the (2, N_out) fp32 mixes are summed over
              NVLink.  The sum is NOT one collective after the last render: the output is cut into
              time segments, and while segment i+1 renders (FP32 pipe) segment i is already being
              reduced (NCCL on its own stream, NVSwitch), so only the last segment's reduction is
              exposed.  Host signals are uploaded time slice by time slice under the same loop.
  by time     given the trajectory, chunks are independent except for the K-1 sample FIR tail
              (apply_hrtf.py:450-453).  The signal is cut at multiples of `chunksize`; each rank
              renders the OUTPUT samples of its segment from its inputs plus a K-1 sample input halo
              (output-stationary: no tail exchange), one MAX all-reduce of the peak implements the
              global normalisation (apply_hrtf.py:462-464), and an all-gather assembles the result.

`local_render` is injectable so the partition / halo / collective logic can be exercised on CPU
ranks (gloo) in tests; the default is the CUDA path and there is no CPU fallback in the product.
"""
from __future__ import annotations

import numpy as np

MIX_SEGMENTS = 6                 # time segments of a by-source mix (render i+1 overlaps the reduction of i)
_SEGMENT_ALIGN = 8192            # output samples: whole render tiles for every tile width


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


def mix_segments(n_out: int, n_segments: int = None):
    """Cut [0, n_out) into about `n_segments` ranges on the render tile grid."""
    n_segments = max(1, int(n_segments or MIX_SEGMENTS))
    step = -(-n_out // n_segments)
    step = (step + _SEGMENT_ALIGN - 1) // _SEGMENT_ALIGN * _SEGMENT_ALIGN
    cuts = list(range(0, n_out, step)) + [n_out]
    return list(zip(cuts[:-1], cuts[1:]))


class PeerMix:
    """Sum of the per-rank mixes over peer memory, fused with the render (csrc/peer.cu): symmetric buffers of
    this rank - a receive block per writer for the slice of the output this rank owns, the full result, and the
    flags - with every peer's mapping of them (torch.distributed._symmetric_memory is the plumbing; all
    arithmetic and all signalling is in libbas_b200.so).  One instance per (group, output length); reusable
    step after step.

        route = peer.route                     -> DeviceRender.render(..., route=route)
        peer.finish(stream)                    -> signal, reduce (rank order: deterministic), wait
        peer.result[:, :n_out]                 -> the full mix on every rank
    """

    def __init__(self, n_out: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _cabi
        self.torch, self._cabi = torch, _cabi
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        device = torch.device('cuda', torch.cuda.current_device())
        n = self.world
        self.n_out = n_out
        self.slice_len = (-(-n_out // n) + _SEGMENT_ALIGN - 1) // _SEGMENT_ALIGN * _SEGMENT_ALIGN
        self.stride = self.slice_len                                # floats between ear rows of a receive block
        self.result_stride = (n_out + 3) // 4 * 4
        recv_floats = n * 2 * self.stride
        result_floats = 2 * self.result_stride
        flag_words = 64 * 3                                         # arrived[n] | done[n] | counter, a cache line apart
        self.buf = symm.empty(recv_floats + result_floats + flag_words, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        torch.cuda.synchronize()
        dist.barrier(group)                                         # every rank's flags are zero before anyone signals
        bases = [int(p) for p in self.handle.buffer_ptrs]
        off_result, off_flags = 4 * recv_floats, 4 * (recv_floats + result_floats)

        def table(offset):
            return torch.tensor([b + offset for b in bases], dtype=torch.int64, device=device)
        self._recv_ptrs = table(0)
        self._result_ptrs = table(off_result)
        self._arrived_ptrs = table(off_flags)
        self._done_ptrs = table(off_flags + 256)
        base = self.buf.data_ptr()
        self._recv, self._arrived, self._done, self._counter = base, base + off_flags, base + off_flags + 256, base + off_flags + 512
        self.result = self.buf[recv_floats:recv_floats + result_floats].view(2, self.result_stride)
        self._own_result_ptr = torch.tensor([bases[self.rank] + off_result], dtype=torch.int64, device=device)
        self.route = _cabi.Route(table_dev=self._recv_ptrs.data_ptr(), n=n, rank=self.rank, len=self.slice_len, stride=self.stride)
        self.epoch = 0
        self._pending_wait = False
        begin = self.rank * self.slice_len
        self.slice = (begin, max(0, min(n_out, begin + self.slice_len) - begin))      # this rank's stretch of the output

    def zero_my_blocks(self):
        """A rank without sources still owes every owner its (all-zero) partial slice."""
        torch = self.torch
        for o in range(self.world):
            block = self.handle.get_buffer(o, (2 * self.stride,), torch.float32, self.rank * 2 * self.stride)
            block.zero_()

    def begin(self, stream):
        """Before a routed render: if the previous step left its result sharded, wait until every owner has finished
        summing that step's slices - only then may this rank's tiles overwrite the owners' receive buffers."""
        if self._pending_wait:
            self._cabi.check(self._cabi.lib.bas_peer_wait(self._done, self.world, self.epoch & 0xffffffff, stream), 'bas_peer_wait')
            self._pending_wait = False

    def finish(self, stream, replicate=True):
        """After this rank's routed render: signal, sum this rank's slice in rank order.  replicate=True: the sum is
        stored into every rank's result buffer and the call waits for all slices - `result` holds the full mix on
        every rank (an all-reduce).  replicate=False: the sum stays here, `result[:, slice]` is this rank's stretch of
        the mix (a reduce-scatter; `gather()` assembles it) and the wait moves to the next `begin`."""
        cabi, lib = self._cabi, self._cabi.lib
        self.epoch += 1
        n, e = self.world, self.epoch & 0xffffffff
        begin, valid = self.slice
        table = self._result_ptrs if replicate else self._own_result_ptr
        cabi.check(lib.bas_peer_reduce(self._recv, n, self.stride, valid, table.data_ptr(), n if replicate else 1, self.result_stride, begin,
                                       self._arrived, e, self._done_ptrs.data_ptr(), self.rank, self._counter,
                                       self._arrived_ptrs.data_ptr(), stream), 'bas_peer_reduce')
        if replicate:
            cabi.check(lib.bas_peer_wait(self._done, n, e, stream), 'bas_peer_wait')
        else:
            self._pending_wait = True

    def gather(self, group=None):
        """The full (2, n_out) mix on every rank from the sharded result (torch all_gather of the slices)."""
        import torch.distributed as dist
        torch = self.torch
        mine = self.result[:, self.rank * self.slice_len:(self.rank + 1) * self.slice_len]
        block = torch.zeros((2, self.slice_len), dtype=torch.float32, device=self.result.device)
        block[:, :mine.shape[1]] = mine
        blocks = [torch.empty_like(block) for _ in range(self.world)]
        dist.all_gather(blocks, block, group=group)
        return torch.cat(blocks, dim=1)[:, :self.n_out]


def _default_render(*args, **kwargs):
    from .apply_hrtf import render_sources
    return render_sources(*args, **kwargs)


_peer_cache = {}


def _peer_mix(n_out, group):
    """PeerMix of (group, output length), built on first use; None where symmetric memory is unavailable."""
    key = (id(group), n_out)
    if key not in _peer_cache:
        try:
            _peer_cache[key] = PeerMix(n_out, group)
        except Exception as e:                                  # no P2P / symmetric memory here: NCCL does the sum
            import warnings
            warnings.warn('peer-memory mix unavailable (%s: %s); using NCCL' % (type(e).__name__, e))
            _peer_cache[key] = None
    return _peer_cache[key]


def render_mix_by_source(signals, chunksize, subchunksize, elev_azim_functions, bank, group=None,
                         dst=None, local_render=None, normalise=True, n_segments=None, exchange='peer'):
    """Every rank passes ONLY its own sources (see shard_sources): an (n_local, N) array (host or CUDA)
    and one trajectory per local source.  Returns the global mix (2, N_out): on every rank when dst is
    None, else only on rank `dst` (other ranks get None).
    exchange='peer' (default): the sum over ranks is fused with the render over NVLink peer memory
    (PeerMix; deterministic rank-order sum); exchange='nccl': one NCCL all_reduce / reduce after the
    last render.  The sum over a rank's own sources is deterministic either way."""
    import torch
    import torch.distributed as dist
    if local_render is not None:
        return _render_mix_injected(torch, dist, signals, chunksize, subchunksize, elev_azim_functions, bank, group, dst,
                                    local_render, normalise)
    from . import apply_hrtf as ah
    from ._cabi import lib, check, decode_status
    torch_dev = torch.device('cuda', torch.cuda.current_device())
    n_local = len(signals)
    rank = dist.get_rank(group)
    # every rank needs the geometry, also one without sources
    n = int(signals.shape[1]) if n_local else 0
    n_t = torch.tensor([n], dtype=torch.int64, device=torch_dev)
    dist.all_reduce(n_t, op=dist.ReduceOp.MAX, group=group)
    n = int(n_t)
    k, n_in, n_out = ah.render_geometry(n, chunksize, subchunksize, bank)
    stride = (n_out + 3) // 4 * 4
    main = torch.cuda.current_stream()
    mix = torch.zeros((2, stride), dtype=torch.float32, device=torch_dev)
    peer = _peer_mix(n_out, group) if exchange == 'peer' and dist.get_world_size(group) > 1 else None
    if exchange == 'peer' and dist.get_world_size(group) > 1:
        # all ranks must take the same path
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=torch_dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not int(ok):
            peer = None
    host_input = n_local and not (isinstance(signals, torch.Tensor) and signals.is_cuda)
    # time segments: a host input is uploaded slice by slice ahead of the segment that needs it; device-resident
    # inputs render in one launch
    segs = mix_segments(n_out, n_segments if (n_segments or host_input) else 1)
    job = None
    uploaded = {}
    if n_local:
        dev = ah._device_bank(bank)
        up, _ = ah._streams(torch, torch_dev)
        if isinstance(signals, torch.Tensor) and signals.is_cuda:
            if signals.dtype == torch.float32 and signals.is_contiguous() and n == n_in:
                x = signals
            else:
                x = torch.zeros((n_local, n_in), dtype=torch.float32, device=torch_dev)
                x[:, :n].copy_(signals)
        else:
            # host signals: uploaded time slice by time slice on a side stream, each slice ahead of the
            # segment that needs it (a segment [pa, pb) reads inputs below pb)
            host = signals.numpy() if isinstance(signals, torch.Tensor) else signals
            host = np.ascontiguousarray(host, dtype=np.float32)
            ah._pin_in_place(host)
            x = ah._scratch(torch, (n_local, n_in), torch.float32, torch_dev)
            up.wait_stream(main)
            if n_in > n:
                with torch.cuda.stream(up):
                    x[:, n:].zero_()                        # zero padding of apply_hrtf.py:405-406
            lo = 0
            for i, (pa, pb) in enumerate(segs):
                hi = min(n, pb)
                if hi > lo:
                    check(lib.bas_copy_2d(x.data_ptr() + 4 * lo, 4 * n_in, host.ctypes.data + 4 * lo, 4 * n, 4 * (hi - lo), n_local, 1,
                                          up.cuda_stream), 'bas_copy_2d')
                    lo = hi
                uploaded[i] = torch.cuda.Event()
                uploaded[i].record(up)
        elev, azim, kinds = ah._directions(elev_azim_functions, n_local, n_in, chunksize)
        n_dirs = n_local * (n_in // chunksize + 1)
        if isinstance(elev, torch.Tensor) or isinstance(azim, torch.Tensor):
            elev_d = torch.as_tensor(elev, dtype=torch.float64).to(torch_dev).contiguous()
            azim_d = torch.as_tensor(azim, dtype=torch.float64).to(torch_dev).contiguous()
        else:
            # The plan kernel reads the directions IN PLACE from pinned host memory (16 B per point over PCIe): a
            # host -> device copy would queue on the copy engine behind the signal upload enqueued above and
            # hold the first render back until the last sample has arrived.
            staged = ah._host_cache.pinned_bytes(torch, 'mix_dirs', 16 * n_dirs)[:16 * n_dirs].view(torch.float64)
            staged[:n_dirs] = torch.from_numpy(np.ascontiguousarray(elev, dtype=np.float64).reshape(-1))
            staged[n_dirs:] = torch.from_numpy(np.ascontiguousarray(azim, dtype=np.float64).reshape(-1))
            elev_d, azim_d = staged[:n_dirs], staged[n_dirs:]
        if elev_d.numel() != n_dirs or azim_d.numel() != n_dirs:
            raise ValueError('trajectories must give %d directions per source' % (n_in // chunksize + 1))
        job = ah.DeviceRender(torch, dev, x, n_in, chunksize, subchunksize, elev_d.reshape(-1), azim_d.reshape(-1), kinds, True,
                              ah._cabi.RENDER_AUTO)
        job.plan(main.cuda_stream)

    def one_pass(gains):
        if peer is not None:
            peer.begin(main.cuda_stream)
        for i, (pa, pb) in enumerate(segs):
            if job is not None:
                if i in uploaded:
                    main.wait_event(uploaded[i])
                job.render(main.cuda_stream, pa, pb, mix.data_ptr() + 4 * pa, stride, gains=gains,
                           route=peer.route if peer is not None else None)
        if peer is not None:
            if job is None:
                peer.zero_my_blocks()
            peer.finish(main.cuda_stream)                   # signal, rank-order sum of this rank's slice, wait for all slices
        elif dst is None:
            dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(mix, dst=dst, op=dist.ReduceOp.SUM, group=group)

    one_pass(None)
    # ---- status, peaks, and the rare second pass of apply_hrtf.py:462-464 (see render_sources) ---------
    # One small MAX all-reduce carries both decisions every rank has to agree on: "some trajectory failed" (then every
    # rank raises - a rank that raised alone would leave the others waiting in the next collective) and "some source
    # peaked above 1" (then every rank runs the second pass).
    flag = torch.zeros(2, dtype=torch.int32, device=torch_dev)
    peaks_host, failure = None, None
    if job is not None:
        small = job.small.cpu().numpy()
        err, where = decode_status(small)
        if err:
            failure = (err, ' (trajectory point %d of local source %d, rank %d)' % (where % job.n_pts, where // job.n_pts, rank))
            flag[1] = 1
        peaks_host = small[2:].view(np.float32)
        if normalise and (peaks_host > 1).any():
            flag[0] = 1
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    second_pass, failed = (int(v) for v in flag.cpu())
    if failed:
        if failure is not None:
            ah._raise_plan_error(*failure)
        raise ah.BasError('a trajectory of another rank failed (see that rank\'s exception)')
    if second_pass:
        gains = None
        if job is not None:
            gains = torch.from_numpy((1.0 / np.maximum(peaks_host, 1.0)).astype(np.float32)).to(torch_dev)
            job.zero_peaks()
        else:
            mix.zero_()
        one_pass(gains)
    result = peer.result[:, :n_out].clone() if peer is not None else mix[:, :n_out]
    if dst is None or rank == dst:
        return result
    return None


def _render_mix_injected(torch, dist, signals, chunksize, subchunksize, elev_azim_functions, bank, group, dst, render, normalise):
    """The same exchange around an injected local renderer (CPU tests): render, then one collective."""
    if len(signals):
        local = render(signals, chunksize, subchunksize, elev_azim_functions, bank, mix=True,
                       normalise=normalise, return_device=True)
    else:
        local = None
    local = _as_tensor(torch, local)
    if local is not None:
        local = local.contiguous()
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
    ranges = time_segments(n_in, chunksize, k, world)
    p0, p1 = ranges[rank]
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
    # all-gather of the time segments: every rank contributes its own (2, longest) block - each output byte
    # crosses NVLink once per receiver - and the blocks are cut back to their true lengths
    longest = max(b - a for a, b in ranges)
    block = torch.zeros((2, longest), dtype=torch.float32, device=dev)
    if seg is not None:
        block[:, :p1 - p0] = seg
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    full = torch.empty((2, n_out), dtype=torch.float32, device=dev)
    for (a, b), blk in zip(ranges, blocks):
        if b > a:
            full[:, a:b] = blk[:, :b - a]
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
    declared = getattr(fn, 'vectorized', None)
    if declared:
        def shifted(t):
            return fn(np.asarray(t) + offset)
        shifted.vectorized = True
        if hasattr(fn, 'az_kind'):
            shifted.az_kind = fn.az_kind
        return shifted
    shifted = lambda t: fn(t + offset)                  # noqa: E731
    if declared is False:
        shifted.vectorized = False
    return shifted


def _as_tensor(torch, value):
    if value is None or isinstance(value, torch.Tensor):
        return value
    return torch.from_numpy(np.ascontiguousarray(value))


def _collective_device(torch, dist, group):
    return torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
