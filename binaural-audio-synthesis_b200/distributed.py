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
    this rank - a receive block per writer for the slice of the output this rank owns, the result, and the
    flags - with every peer's mapping of them (torch.distributed._symmetric_memory is the plumbing; all
    arithmetic and all signalling is in libbas_b200.so).  One instance per (group, output length); reusable
    step after step.  Receive and result buffers exist DEPTH times (step number modulo DEPTH), for the pipelined form.

    One step at a time:
        peer.begin(stream); DeviceRender.render(..., route=peer.route); peer.finish(stream)
        peer.result[:, :n_out]                 -> the full mix on every rank (replicate=True), or this rank's slice

    Pipelined, for a stream of steps (batches): the exchange of step i runs on a side stream while the main stream
    plans and renders steps i + 1 and i + 2 into the other receive buffers (a persistent render fills every SM, so the
    reduce kernel of step i usually runs between two renders, beside the plan kernel; the waits for the peers are
    stream memory operations and hold no SM) -
        route = peer.submit_route()            -> DeviceRender.render(..., route=route): the render's last CTA signals
        peer.submit(stream, replicate=False)   -> side stream: wait for all writers, rank-order sum, wait for all owners
        ...                                       (the next steps)
        peer.flush(stream)                     -> `stream` waits for every submitted exchange; peer.result = the latest
    """

    DEPTH = 3            # sets of receive / result buffers: the exchange of step i may still run while step i + 2 renders

    def __init__(self, n_out: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _cabi
        self.torch, self._cabi = torch, _cabi
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        device = torch.device('cuda', torch.cuda.current_device())
        self.device = device
        n = self.world
        self.n_out = n_out
        self.slice_len = (-(-n_out // n) + _SEGMENT_ALIGN - 1) // _SEGMENT_ALIGN * _SEGMENT_ALIGN
        self.stride = self.slice_len                                # floats between ear rows of a receive block
        self.result_stride = (n_out + 3) // 4 * 4
        recv_floats = n * 2 * self.stride
        result_floats = 2 * self.result_stride
        flag_words = 64 * 4                      # arrived[n] | done[n] | reduce counter | render counter, a cache line apart
        d = self.DEPTH
        self.buf = symm.empty(d * (recv_floats + result_floats) + flag_words, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        torch.cuda.synchronize()
        dist.barrier(group)                                         # every rank's flags are zero before anyone signals
        bases = [int(p) for p in self.handle.buffer_ptrs]
        self._recv_floats = recv_floats
        off_recv = [4 * recv_floats * k for k in range(d)]
        off_result = [4 * (d * recv_floats + result_floats * k) for k in range(d)]
        off_flags = 4 * d * (recv_floats + result_floats)

        def table(offset):
            return torch.tensor([b + offset for b in bases], dtype=torch.int64, device=device)
        self._recv_ptrs = [table(o) for o in off_recv]
        self._result_ptrs = [table(o) for o in off_result]
        self._arrived_ptrs = table(off_flags)
        self._done_ptrs = table(off_flags + 256)
        base = self.buf.data_ptr()
        self._recv = [base + o for o in off_recv]
        self._arrived, self._done = base + off_flags, base + off_flags + 256
        self._counter, self._render_counter = base + off_flags + 512, base + off_flags + 768
        self.results = [self.buf[d * recv_floats + result_floats * k:d * recv_floats + result_floats * (k + 1)].view(2, self.result_stride)
                        for k in range(d)]
        self.result = self.results[0]
        self._own_result_ptr = [torch.tensor([bases[self.rank] + o], dtype=torch.int64, device=device) for o in off_result]
        self._routes = [_cabi.Route(table_dev=self._recv_ptrs[k].data_ptr(), n=n, rank=self.rank, len=self.slice_len, stride=self.stride)
                        for k in range(d)]
        self._signal_routes = [_cabi.Route(table_dev=self._recv_ptrs[k].data_ptr(), n=n, rank=self.rank, len=self.slice_len, stride=self.stride,
                                           arrive_ptrs_dev=self._arrived_ptrs.data_ptr(), arrive_counter_dev=self._render_counter)
                               for k in range(d)]
        self.route = self._routes[0]
        self.epoch = 0
        self._latest = 0
        self._pending_wait = False
        self._stream_wait_ok = True
        self._side = torch.cuda.Stream(device=device, priority=-1)
        self._rendered = torch.cuda.Event()
        self._exchanged = [None] * d                                  # per parity: event "exchange of the last step of this parity complete"
        begin = self.rank * self.slice_len
        self.slice = (begin, max(0, min(n_out, begin + self.slice_len) - begin))      # this rank's stretch of the output

    def _stream(self, handle):
        """torch's view of a raw stream handle.  Handle 0 is the default stream: ExternalStream(0) would NOT wrap it
        (torch treats a null pointer as 'no stream given' and takes a fresh one from its pool)."""
        torch = self.torch
        return torch.cuda.default_stream(self.device) if not handle else torch.cuda.ExternalStream(handle, device=self.device)

    def zero_my_blocks(self, parity=0):
        """A rank without sources still owes every owner its (all-zero) partial slice."""
        torch = self.torch
        for o in range(self.world):
            block = self.handle.get_buffer(o, (2 * self.stride,), torch.float32, parity * self._recv_floats + self.rank * 2 * self.stride)
            block.zero_()

    def begin(self, stream):
        """Before a routed render: if the previous step left its result sharded, wait until every owner has finished
        summing that step's slices - only then may this rank's tiles overwrite the owners' receive buffers."""
        self.flush(stream)
        if self._pending_wait:
            self._cabi.check(self._cabi.lib.bas_peer_wait(self._done, self.world, self.epoch & 0xffffffff, stream), 'bas_peer_wait')
            self._pending_wait = False

    def _reduce(self, parity, e, replicate, fold_signal, stream):
        cabi, lib = self._cabi, self._cabi.lib
        n = self.world
        begin, valid = self.slice
        table = self._result_ptrs[parity] if replicate else self._own_result_ptr[parity]
        cabi.check(lib.bas_peer_reduce(self._recv[parity], n, self.stride, valid, table.data_ptr(), n if replicate else 1, self.result_stride, begin,
                                       self._arrived, e, self._done_ptrs.data_ptr(), self.rank, self._counter,
                                       self._arrived_ptrs.data_ptr() if fold_signal else None, stream), 'bas_peer_reduce')

    def finish(self, stream, replicate=True):
        """After this rank's routed render: signal, sum this rank's slice in rank order.  replicate=True: the sum is
        stored into every rank's result buffer and the call waits for all slices - `result` holds the full mix on
        every rank (an all-reduce).  replicate=False: the sum stays here, `result[:, slice]` is this rank's stretch of
        the mix (a reduce-scatter; `gather()` assembles it) and the wait moves to the next `begin`."""
        self.epoch += 1
        e = self.epoch & 0xffffffff
        self._reduce(0, e, replicate, True, stream)
        self.result = self.results[0]
        if replicate:
            self._cabi.check(self._cabi.lib.bas_peer_wait(self._done, self.world, e, stream), 'bas_peer_wait')
        else:
            self._pending_wait = True

    # ---- pipelined form ----------------------------------------------------------------------------
    def submit_route(self, stream, signal=True):
        """The route of the step about to be rendered on `stream` (cuda stream handle): the receive buffers of its
        parity, which `stream` may write once every owner has summed the step that used them last.  signal=True (the
        step's LAST routed render): the render kernel itself tells the owners when this rank's tiles have landed."""
        if self._pending_wait:                                       # a sharded step of the one-at-a-time form came before
            self.begin(stream)
        parity = (self.epoch + 1) % self.DEPTH
        ev = self._exchanged[parity]
        if ev is not None:
            self._stream(stream).wait_event(ev)
            self._exchanged[parity] = None
        if not signal:
            return self._routes[parity]
        route = self._signal_routes[parity]
        route.arrive_epoch = (self.epoch + 1) & 0xffffffff
        return route

    def submit(self, stream, replicate=False, rendered=True):
        """After the step's routed renders on `stream`: its exchange goes to the side stream - wait for every writer, sum
        this rank's slice in rank order into results[parity] (of every rank: replicate), wait until every owner is
        done.  `stream` is not held up; flush() or the submit_route() DEPTH steps later joins it."""
        torch, cabi, lib = self.torch, self._cabi, self._cabi.lib
        self.epoch += 1
        e = self.epoch & 0xffffffff
        parity = self.epoch % self.DEPTH
        if not rendered:                                            # no sources here: zero blocks, explicit signal
            with torch.cuda.stream(self._stream(stream)):
                self.zero_my_blocks(parity)
            cabi.check(lib.bas_peer_signal(self._arrived_ptrs.data_ptr(), self.world, self.rank, e, stream), 'bas_peer_signal')
        self._rendered.record(self._stream(stream))
        self._side.wait_event(self._rendered)
        side = self._side.cuda_stream
        # both waits as stream memory operations where the driver offers them: no CTA sits on an SM while a peer is late
        # (the reduce kernel's own check of the arrival flags then passes at once)
        self._side_wait(self._arrived, e)
        self._reduce(parity, e, replicate, False, side)
        self._side_wait(self._done, e)
        ev = torch.cuda.Event()
        ev.record(self._side)
        self._exchanged[parity] = ev
        self._latest = parity
        return parity, ev

    def _side_wait(self, flags, e):
        cabi, lib = self._cabi, self._cabi.lib
        side = self._side.cuda_stream
        if self._stream_wait_ok:
            rc = lib.bas_peer_stream_wait(flags, self.world, e, side)
            if rc == 0:
                return
            if rc != cabi.E_UNSUPPORTED:
                cabi.check(rc, 'bas_peer_stream_wait')
            self._stream_wait_ok = False
        cabi.check(lib.bas_peer_wait(flags, self.world, e, side), 'bas_peer_wait')

    def collect(self, ticket, stream):
        """The result buffer (2, result_stride) of the step submit() returned `ticket` for, once `stream` has waited for
        its exchange.  Valid until DEPTH - 1 further steps have been submitted."""
        parity, ev = ticket
        self._stream(stream).wait_event(ev)
        return self.results[parity]

    def flush(self, stream):
        """`stream` waits for every submitted exchange; `result` is then the mix of the last submitted step."""
        ext = None
        for k in range(self.DEPTH):
            if self._exchanged[k] is not None:
                ext = ext or self._stream(stream)
                ext.wait_event(self._exchanged[k])
                self._exchanged[k] = None
        if ext is not None:
            self.result = self.results[self._latest]

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


def render_mix_stream(batches, chunksize, subchunksize, bank, group=None, replicate=True):
    """A STREAM of batches of sources, each mixed over the ranks of `group` like render_mix_by_source(..., normalise=False),
    the exchange of a batch running beside the plan and render of the next one (PeerMix, pipelined form).

    batches: iterable of (signals, elev_azim_functions) - this rank's share of every batch: a CUDA float32 tensor
    (n_local >= 1, N), the same shape for every batch, and its trajectories (as render_sources takes them).
    Yields one (2, N_out) CUDA tensor per batch, in order, each after the NEXT batch has been enqueued (the last one
    after the loop).  replicate=True: the full mix on every rank; False: a reduce-scatter - only the columns
    [rank * slice_len, (rank + 1) * slice_len) hold this rank's stretch of the sum.  The tensor is a view of the
    exchange buffers, valid until two more batches have been drawn from the generator: copy it to keep it.  No normalisation
    (apply_hrtf.py:462-464 is per source).  A trajectory error on any rank raises on every rank."""
    import torch
    import torch.distributed as dist
    from . import apply_hrtf as ah
    from ._cabi import decode_status, RENDER_AUTO
    device = torch.device('cuda', torch.cuda.current_device())
    world = dist.get_world_size(group)
    main = torch.cuda.current_stream()
    st = main.cuda_stream
    dev = ah._device_bank(bank)
    state = {'peer': None, 'asked': False, 'scratch': None}

    def enqueue(signals, trajs):
        if not (isinstance(signals, torch.Tensor) and signals.is_cuda and signals.dtype == torch.float32 and signals.dim() == 2 and len(signals)):
            raise ValueError('render_mix_stream takes CUDA float32 signals (n_local >= 1, N)')
        n_local, n = signals.shape
        k, n_in, n_out = ah.render_geometry(n, chunksize, subchunksize, bank)
        stride = (n_out + 3) // 4 * 4
        if signals.is_contiguous() and n == n_in:
            x = signals
        else:
            x = torch.zeros((n_local, n_in), dtype=torch.float32, device=device)
            x[:, :n].copy_(signals)
        elev, azim, kinds = ah._directions(trajs, n_local, n_in, chunksize)
        elev_d = torch.as_tensor(elev, dtype=torch.float64).to(device).contiguous().reshape(-1)
        azim_d = torch.as_tensor(azim, dtype=torch.float64).to(device).contiguous().reshape(-1)
        if elev_d.numel() != n_local * (n_in // chunksize + 1) or azim_d.numel() != elev_d.numel():
            raise ValueError('trajectories must give %d directions per source' % (n_in // chunksize + 1))
        job = ah.DeviceRender(torch, dev, x, n_in, chunksize, subchunksize, elev_d, azim_d, kinds, True, RENDER_AUTO)
        job.plan(st)
        if world > 1 and not state['asked']:
            peer = _peer_mix(n_out, group)
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)           # all ranks take the same path
            state['peer'], state['asked'] = (peer if int(ok) else None), True
        peer = state['peer']
        if peer is not None:
            if state['scratch'] is None:
                state['scratch'] = torch.zeros((2, stride), dtype=torch.float32, device=device)     # never written: every tile has an owner
            job.render(st, 0, n_out, state['scratch'].data_ptr(), stride, route=peer.submit_route(st))
            ticket = peer.submit(st, replicate=replicate)
            out = None
        else:
            out = torch.zeros((2, stride), dtype=torch.float32, device=device)
            job.render(st, 0, n_out, out.data_ptr(), stride)
            ticket = None
        flag = (job.small[:1] != 0).to(torch.int32)
        work = None
        if world > 1:
            if peer is None:
                dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
            work = dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group, async_op=True)
        return job, ticket, out, flag, work, n_out

    def collect(entry):
        job, ticket, out, flag, work, n_out = entry
        if work is not None:
            work.wait()
        if int(flag):                                                # the batch is a step old: this wait costs the device nothing
            bits, index = decode_status(job.small.cpu().numpy())
            if bits:
                ah._raise_plan_error(bits, ' (trajectory point %d of source %d)' % (index % job.n_pts, index // job.n_pts))
            raise ah.BasError('render_mix_stream: a trajectory failed on another rank')
        if ticket is not None:
            out = state['peer'].collect(ticket, st)
        return out[:, :n_out]

    pending = None
    for signals, trajs in batches:
        entry = enqueue(signals, trajs)
        if pending is not None:
            yield collect(pending)
        pending = entry
    if pending is not None:
        yield collect(pending)


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
