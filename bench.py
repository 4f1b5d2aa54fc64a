#!/usr/bin/env python
"""Benchmark of the moving-source binaural render path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

Under torchrun (N > 1) one rank per GPU over NCCL.  --config selects the workload:

  mix64 (default)  BASELINE.json configs[2]: 64 independent 60 s sources, each on its own Lissajous
                   trajectory, mixed to ONE binaural output; sources sharded round-robin over the
                   ranks (strong scaling: the job is the same at every N) and the per-rank mixes
                   summed INSIDE the timed region: fused with the render kernel over NVLink peer memory
                   (--collective peer_pipelined, the default: the exchange of a step runs beside the plan
                   and render of the next, the timed region ends when the last exchange is complete
                   everywhere, summed mix left sharded by time; peer_sharded: the exchange inside its own
                   step; peer: replicated on every rank) or by NCCL (all_reduce / reduce, optionally per
                   time segment).  A step = plan (64 x 5169 directions) + render/mix + exchange.
  single           configs[1]: one 60 s source per rank (weak scaling, no collective).
  hour             configs[3]: one 1-hour 48 kHz source cut across the ranks by time (K-1 halo),
                   global peak by MAX all-reduce.
  stress1024       configs[4]: 1024 sources x 60 s, full-length IRs (K = 512), N = 16 bank.

value      whole-job source-sample-pairs/s (for one source: output sample-pairs/s), inputs resident in
           HBM, CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
e2e        the same job through the public API with HOST arrays: pageable ndarrays and plain-lambda
           trajectories in (what a caller of the reference passes), host array out, copies inside the
           timed region.  e2e_fast: the opt-in form (pinned arrays, trajectories declared vectorised).
roofline   the dominant kernel of the step (the mixing render kernel): algorithmic HBM bytes
           ((4 n_src + 8) B per output pair) over its CUDA-event duration against MEASURED_PEAKS.json,
           plus - because the path is bound by the FP32 pipe - useful FMA/s against the
           measured and the nominal FMA peak.  single_source carries configs[1]'s figures.
parity     N-GPU mix against rank 0 rendering all sources alone (window), and against the sum of
           per-source renders through the non-mixing kernel.
cpu_baseline / --impl reference   the UNMODIFIED reference (oracle/_ref bytecode, kind "reference") or,
           where that is not built, the numpy port (kind "port"), on the host cores.
"""
import argparse
import contextlib
import importlib.util
import io
import json
import os
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK, SUB = 512, 32
METRIC = 'binaural output sample-pairs/s'
UNIT = 'sample-pairs/s'
NOMINAL_TFMA = 148 * 128 * 1.965e9 / 1e12            # 37.2: 148 SMs x 128 FP32 lanes x 1.965 GHz

CONFIGS = {
    # name: (BASELINE.json index, sources, seconds, fs, samples_to_keep, upsampling, sigma)
    'single': (1, 1, 60, 44100, 256, 8, 0.05),
    'mix64': (2, 64, 60, 44100, 256, 8, 0.05 / 8),
    'hour': (3, 1, 3600, 48000, 256, 8, 0.05),
    'stress1024': (4, 1024, 60, 44100, 512, 16, 0.05 / 32),
}
ALIASES = {'2': 'single', '3': 'mix64', '4': 'hour', '5': 'stress1024'}      # SURVEY.md's 1-based config numbers


def workload_name(name):
    idx, n_src, secs, fs, keep, ups, _ = CONFIGS[name]
    what = {'single': 'one %d s mono source on a Lissajous az+el trajectory' % secs,
            'mix64': '%d independent %d s sources, each on its own Lissajous trajectory, mixed to one binaural output' % (n_src, secs),
            'hour': 'one %d s source time-segmented across the GPUs with a K-1 halo' % secs,
            'stress1024': '%d sources x %d s mixed to one binaural output, full-length IRs' % (n_src, secs)}[name]
    return 'configs[%d]: %s; %g kHz, N=%d bank, samples_to_keep=%d, chunksize=%d, subchunksize=%d' % (
        idx, what, fs / 1e3, ups, keep, CHUNK, SUB)


# ---------------------------------------------------------------------------------------------------
# synthetic inputs (no product library needed: the reference arm uses these too)
# ---------------------------------------------------------------------------------------------------
def load_bank_synth():
    """bank_synth.py without importing the package (which would load libbas_b200.so)."""
    pkg_dir = os.path.join(ROOT, 'binaural-audio-synthesis_b200')
    name = '_bas_nolib'
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [pkg_dir]
        sys.modules[name] = pkg
    spec = importlib.util.spec_from_file_location(name + '.bank_synth', os.path.join(pkg_dir, 'bank_synth.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name + '.bank_synth'] = mod
    spec.loader.exec_module(mod)
    return mod


def make_bank(ups=8, keep=256):
    if not isinstance(ups, int):                   # tools/ written against round 1 pass the package here
        ups, keep = 8, 256
    f = load_bank_synth().build_bank(ups, seed=0)

    class bank:
        upsampling = ups
        diffs_left, diffs_right = f['diffs_left'], f['diffs_right']
        irs_left, irs_right = f['irs_left'][:, :keep * ups], f['irs_right'][:, :keep * ups]
    return bank


def lissajous(s, fs=44100):
    """Trajectory of source s: elevation in [-45, 90] deg crossing every ring, azimuth circling.  Plain
    lambda, arithmetic on t only - like the reference's own (apply_hrtf.py:583-593)."""
    rng = np.random.default_rng(3000 + s)
    f1, f2 = (3.0, 5.0) if s == 0 else (rng.uniform(2, 4), rng.uniform(3, 6))
    p1, p2 = (0.3, 1.0) if s == 0 else (rng.uniform(0, 6), rng.uniform(0, 6))
    k = np.float64(2 * np.pi / (4 * fs))
    return lambda t: (np.deg2rad(22.5 + 67.5 * np.sin(f1 * k * t + p1)), (f2 * k * t + p2) % (2 * np.pi))


def spiral(fs, length):
    """apply_hrtf.py:590-593 with length = the signal's: a slow spiral from -45 deg to the pole."""
    turns = length / 8.0
    return lambda t: ((-np.pi / 4) + (3 * np.pi / 4) * (t / (fs * length)), 2 * np.pi * t * turns / (fs * length))


def pink_noise(n, seed):
    """1/f-shaped Gaussian noise, sigma = 0.05 (configs[1]'s signal; used by tools/)."""
    rng = np.random.default_rng(seed)
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.arange(spec.size, dtype=np.float64)
    f[0] = 1.0
    x = np.fft.irfft(spec / np.sqrt(f), n)
    return (0.05 * x / x.std()).astype(np.float32)


def noise_host(n, seed, sigma):
    return (sigma * np.random.default_rng(seed).standard_normal(n, dtype=np.float32)).astype(np.float32)


def pin_to_device_cpus(index):
    """Run this rank on the CPU cores NVML names as local to GPU `index` (same NUMA node / PCIe root)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return 'rank pinned to the %d CPU cores local to its GPU (NVML)' % len(cpus)
    except Exception as e:
        return 'cpu affinity not set (%s)' % type(e).__name__
    return None


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs (NVML, ~2 ms period)."""

    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap',
               0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on all host cores
# ---------------------------------------------------------------------------------------------------
def cpu_renderer():
    """(make_signal_move_2d, Bank constructor, kind): the unmodified reference from oracle/_ref when it was
    built (kind 'reference'), else the numpy port of oracle/ (kind 'port')."""
    from oracle import reference_loader
    if reference_loader.available():
        ref, _ = reference_loader.load()

        def render(x, c, s, traj, bank):
            with contextlib.redirect_stdout(io.StringIO()):          # the reference prints progress (apply_hrtf.py:456)
                return ref.make_signal_move_2d(x, c, s, traj, bank)
        return render, 'reference'
    from oracle import binaural_oracle as oracle
    return oracle.make_signal_move_2d, 'port'


def _bank_object(fields):
    class bank:
        upsampling, diffs_left, diffs_right, irs_left, irs_right = fields
    return bank


_CPU = {}


def _cpu_stretch(args):
    """One process: one stretch of one source, rendered by the reference."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    if 'render' not in _CPU:
        _CPU['render'], _CPU['kind'] = cpu_renderer()
    x, t0, fields, s, fs, hour = args
    traj = spiral(fs, 3600) if hour else lissajous(s, fs)
    out = _CPU['render'](x, CHUNK, SUB, lambda t: traj(t + t0), _bank_object(fields))
    return out.shape[0]


def run_reference(args, rank, world):
    """bench.py --impl reference: the reference's own make_signal_move_2d (apply_hrtf.py:356-466) on every host
    core.  A step renders `cores` independent 1.5 s stretches of the workload's sources, one per process -
    a bounded sample of the same job (the mix itself is one addition per sample on top)."""
    if rank != 0:
        return
    import multiprocessing as mp
    name = args.config
    idx, n_src, secs, fs, keep, ups, sigma = CONFIGS[name]
    bank = make_bank(ups, keep)
    fields = (ups, bank.diffs_left, bank.diffs_right, bank.irs_left, bank.irs_right)
    _, kind = cpu_renderer()
    cores = os.cpu_count() or 1
    seg = int(1.5 * fs) // CHUNK * CHUNK
    n = secs * fs
    jobs = []
    for i in range(cores):
        s = i % n_src
        start = (i * 7919 * CHUNK) % (n - seg) // CHUNK * CHUNK
        jobs.append((noise_host(seg, 100 + i, sigma), start, fields, s, fs, name == 'hour'))
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(_cpu_stretch, jobs)
        t0 = time.perf_counter()
        pairs = 0
        for _ in range(args.steps):
            pairs += sum(pool.map(_cpu_stretch, jobs))
        dt = time.perf_counter() - t0
    value = pairs / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak' if name == 'single' else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(name), 'parallelism': 'host processes, one per stretch'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': '%d stretches of %.2f s of the workload\'s sources per step, one per process (%s)' % (
                             cores, seg / fs, 'unmodified reference, oracle/_ref' if kind == 'reference' else 'numpy port, oracle/')},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class Ctx:
    """What every leg of our arm needs."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import binaural_audio_synthesis_b200 as bas
        self.torch, self.bas, self.cabi, self.lib = torch, bas, bas._cabi, bas._cabi.lib
        self.args, self.rank, self.world = args, rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device('cuda', local_rank)
        self.local_rank = local_rank
        self.cpu_note = pin_to_device_cpus(local_rank) if world > 1 and not os.environ.get('BAS_NO_CPU_AFFINITY') else None
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=self.dev)
            self.dist = dist
        bas.apply_hrtf.PROGRESS = False
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def gather_objects(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def timed(self, fn, steps, warmup, collective=True, after=None):
        """(ms per step, per-step ms list) of fn(i) over `steps` steps after `warmup`: CUDA events on the
        launching stream, a barrier + synchronize on both sides, max over ranks.  `after` (joins work the steps left
        on other streams) runs inside the timed region, before the closing event."""
        torch = self.torch
        sync = self.barrier if collective else torch.cuda.synchronize
        for i in range(warmup):
            fn(i)
        if after:
            after()
        sync()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record(self.stream)
        for i in range(steps):
            fn(warmup + i)
            if after and i == steps - 1:
                after()
            marks[i + 1].record(self.stream)
        sync()
        total = marks[0].elapsed_time(marks[-1])
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        if collective:
            total = self.max_over_ranks(total)
        return total / steps, per_step

    def device_noise(self, n, seed, sigma, n_pad=None):
        """White Gaussian noise generated on the device (synthetic input, not part of any timed region);
        the same seed gives the same samples on every B200 of the box."""
        torch = self.torch
        g = torch.Generator(device=self.dev)
        g.manual_seed(int(seed))
        x = torch.zeros(n_pad or n, dtype=torch.float32, device=self.dev)
        x[:n] = torch.randn(n, generator=g, device=self.dev, dtype=torch.float32) * sigma
        return x

    def hbm_peak(self):
        path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(path):
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        return 6650.0, 'fallback (B200_PROFILING.md)'

    def fma_probe(self):
        """FP32 FMA peak measured here (libbas_probe.so, not the product library)."""
        torch = self.torch
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import probe_lib
        probe = probe_lib.load()
        sink = torch.empty(148 * 8 * 256, dtype=torch.float32, device=self.dev)
        st = self.stream.cuda_stream
        out = {}
        for name, packed in (('fma_f32', 0), ('fma_f32x2', 1)):
            iters = 4096
            ms, _ = self.timed(lambda i: probe.bas_probe_fma(packed, 148 * 8, 256, iters, sink.data_ptr(), st), 5, 2, collective=False)
            out[name] = 148 * 8 * 256 * iters * 32 / (ms * 1e-3) / 1e12
        out['nominal'] = NOMINAL_TFMA
        return out


def traffic_of(mix):
    """DRAM bytes per launch of the render kernel (mixing or one source per tile) from the committed ncu --set full
    capture (profiles/r2_traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if not os.path.exists(path):
        return None, None
    tj = json.load(open(path))
    for name, kv in tj['kernels'].items():
        if name.startswith('bas_render_tiled_kernel') and bool(kv.get('mix')) == bool(mix):
            return kv['dram_bytes_read'] + kv['dram_bytes_write'], 'profiles/r2_traffic.json (%s): %s, %s' % (kv['report'], name, kv['command'])
    return None, None


def single_source_block(ctx, bank, fs, secs, sigma, steps):
    """configs[1]: one 60 s source at a time on this GPU - the step (plan -> render with fused filter
    synthesis -> normalise, one bas_render_step call), the render kernel alone against both roofs, and the
    public call make_signal_move_2d with host arrays."""
    torch, bas, cabi, lib = ctx.torch, ctx.bas, ctx.cabi, ctx.lib
    ah = bas.apply_hrtf
    bdev = ah._device_bank(bank)
    n = secs * fs
    k, n_in, n_out = bas.render_geometry(n, CHUNK, SUB, bank)
    n_pts = n_in // CHUNK + 1
    times = np.arange(0, n_in + 1, CHUNK, dtype=np.int64)
    out_stride = (n_out + 3) // 4 * 4
    n_sets = 8                                   # rotating over more bytes than L2 holds
    sets = []
    for i in range(n_sets):
        x = ctx.device_noise(n, 7000 + i, sigma, n_in)[None, :]
        elev, azim = lissajous(i, fs)(times)
        job = ah.DeviceRender(torch, bdev, x, n_in, CHUNK, SUB, torch.from_numpy(np.ascontiguousarray(elev)).to(ctx.dev),
                              torch.from_numpy(np.ascontiguousarray(azim)).to(ctx.dev), cabi.AZ_F64, False, ctx.args.variant)
        job.workspace = torch.empty(int(lib.bas_render_workspace_bytes()), dtype=torch.uint8, device=ctx.dev)
        job.job.workspace_dev = job.workspace.data_ptr()
        sets.append((job, torch.empty((1, 2, out_stride), dtype=torch.float32, device=ctx.dev)))
    st = ctx.stream.cuda_stream

    def step(i, stream=None):
        job, out = sets[i % n_sets]
        j = job.job
        j.flags = cabi.STEP_PLAN | cabi.STEP_RENDER | cabi.STEP_NORMALISE | job._fused_flag
        j.p_begin, j.p_count, j.out_dev, j.out_stride = 0, n_out, out.data_ptr(), out_stride
        cabi.check(lib.bas_render_step(j, stream or st), 'bas_render_step')

    def render_only(i):
        job, out = sets[i % n_sets]
        job.render(st, 0, n_out, out.data_ptr(), out_stride)

    for i in range(n_sets):
        step(i)                                   # every set planned once (render_only needs the terms)
    ms_serial, _ = ctx.timed(step, steps, 5, collective=False)
    # several independent sources in flight: consecutive steps alternate over CUDA streams, the way a server
    # keeps several make_signal_move_2d calls going on one GPU
    n_flight = 4
    streams = [ctx.stream] + [torch.cuda.Stream() for _ in range(n_flight - 1)]

    def flight(i):
        lane = i % n_flight
        if lane:
            streams[lane].wait_stream(ctx.stream) if i < n_flight else None
        step(i, streams[lane].cuda_stream)
    torch.cuda.synchronize()
    for i in range(8):
        flight(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for lane in range(1, n_flight):
        streams[lane].wait_event(e0)
    for i in range(steps):
        step(i, streams[i % n_flight].cuda_stream)
    for lane in range(1, n_flight):
        ev = torch.cuda.Event()
        ev.record(streams[lane])
        ctx.stream.wait_event(ev)
    e1.record(ctx.stream)
    torch.cuda.synchronize()
    ms_flight = e0.elapsed_time(e1) / steps
    ms_render, _ = ctx.timed(render_only, max(steps, 20), 3, collective=False)
    assert cabi.decode_status(sets[0][0].small.cpu().numpy())[0] == 0
    hbm_peak, peak_src = ctx.hbm_peak()
    algo = 12.0 * n_out
    useful_fma = 2.0 * k * n_in
    fma = ctx.fma_probe()
    traffic, traffic_src = traffic_of(False)
    block = {
        'workload': workload_name('single'),
        'ms_per_step_one_at_a_time': ms_serial, 'ms_per_step_4_in_flight': ms_flight,
        'value_one_at_a_time': n_out / (ms_serial * 1e-3), 'value_4_in_flight': n_out / (ms_flight * 1e-3),
        'fused_filter_synthesis': bool(sets[0][0].fused),
        'kernels_per_step': 'memset, plan_build, render (filter rows synthesised in-kernel), normalise - one bas_render_step call',
        'roofline': {'bound': 'hbm', 'kernel': 'bas_render_tiled_kernel (one source, fused filter synthesis)', 'achieved': algo / (ms_render * 1e-3) / 1e9,
                     'peak': hbm_peak, 'unit': 'GB/s', 'frac': algo / (ms_render * 1e-3) / 1e9 / hbm_peak, 'ms_per_launch': ms_render,
                     'algorithmic_bytes_per_launch': algo, 'peak_source': peak_src, 'traffic': traffic, 'traffic_source': traffic_src,
                     'fp32_pipe': {'achieved_tfma_s': useful_fma / (ms_render * 1e-3) / 1e12, 'peak_measured_tfma_s': max(fma['fma_f32'], fma['fma_f32x2']),
                                   'peak_nominal_tfma_s': NOMINAL_TFMA, 'frac_of_nominal': useful_fma / (ms_render * 1e-3) / 1e12 / NOMINAL_TFMA,
                                   'probe': fma, 'note': 'useful FMAs only: 2 ears x K taps per input sample'}},
    }
    # ---- the public call with host arrays -----------------------------------------------------------
    x_page = noise_host(n, 7000, sigma)                           # an ordinary (pageable) ndarray
    traj = lissajous(0, fs)                                       # a plain lambda
    for _ in range(4):
        y = bas.make_signal_move_2d(x_page, CHUNK, SUB, traj, bank)
    torch.cuda.synchronize()
    reps = max(5, min(steps, 30))
    t0 = time.perf_counter()
    for _ in range(reps):
        y = bas.make_signal_move_2d(x_page, CHUNK, SUB, traj, bank)
    dt = (time.perf_counter() - t0) / reps
    assert y.shape == (n_out, 2)
    fresh = noise_host(n, 7001, sigma)
    t0 = time.perf_counter()
    bas.make_signal_move_2d(fresh, CHUNK, SUB, traj, bank)
    first_ms = 1e3 * (time.perf_counter() - t0)
    x_pin = torch.from_numpy(x_page.copy()).pin_memory().numpy()
    fast = lambda t: traj(t)                                      # noqa: E731
    fast.vectorized = True
    for _ in range(3):
        bas.make_signal_move_2d(x_pin, CHUNK, SUB, fast, bank)
    t0 = time.perf_counter()
    for _ in range(reps):
        bas.make_signal_move_2d(x_pin, CHUNK, SUB, fast, bank)
    dt_fast = (time.perf_counter() - t0) / reps
    block['e2e'] = {'value': n_out / dt, 'ms_per_call': 1e3 * dt, 'unit': UNIT,
                    'call': 'make_signal_move_2d(pageable float32 ndarray, 512, 32, plain lambda, bank) -> host ndarray; the array is '
                            'page-locked in place on its first call (first_call_ms, a fresh array) and re-used while it lives',
                    'first_call_ms': first_ms, 'h2d_bytes': int(4 * n + 16 * n_pts), 'd2h_bytes': int(8 * n_out + 12)}
    block['e2e_fast'] = {'value': n_out / dt_fast, 'ms_per_call': 1e3 * dt_fast, 'unit': UNIT,
                         'call': 'the opt-in form: input already in pinned memory, trajectory declared vectorised'}
    return block


def run_mix(ctx, name):
    """mix64 / stress1024 / single: sources sharded round-robin; per-rank mix; NCCL sum per time segment."""
    args, torch, bas, cabi, lib, dist = ctx.args, ctx.torch, ctx.bas, ctx.cabi, ctx.lib, ctx.dist
    ah = bas.apply_hrtf
    from binaural_audio_synthesis_b200 import distributed
    idx, n_src, secs, fs, keep, ups, sigma = CONFIGS[name]
    rank, world = ctx.rank, ctx.world
    weak = name == 'single'
    if weak:
        n_src = world                                             # one source per rank
    bank = make_bank(ups, keep)
    bdev = ah._device_bank(bank)
    n = secs * fs
    k, n_in, n_out = bas.render_geometry(n, CHUNK, SUB, bank)
    n_pts = n_in // CHUNK + 1
    stride = (n_out + 3) // 4 * 4
    times = np.arange(0, n_in + 1, CHUNK, dtype=np.int64)
    mine = distributed.shard_sources(n_src, rank, world)
    n_local = len(mine)
    mix_mode = not weak
    collective = args.collective if (world > 1 and mix_mode) else 'none'
    peer = None
    sharded = collective in ('peer_sharded', 'peer_pipelined')
    pipelined = collective == 'peer_pipelined'
    if collective in ('peer', 'peer_sharded', 'peer_pipelined'):
        peer = distributed._peer_mix(n_out, None)
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=ctx.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok):
            peer, collective = None, 'all_reduce'
    # device-resident inputs: one render launch per step (the exchange needs no time segments when it is fused
    # with the render); --segments N cuts the output into N launches, each followed by its NCCL collective
    segs = distributed.mix_segments(n_out, args.segments or 1) if mix_mode else [(0, n_out)]

    # ---- inputs resident in HBM: signals, directions.  Rotating sets so that a step's working set is not
    #      L2-resident from the step before (the 126 MB L2 would otherwise serve the signals) --------------
    set_bytes = 4 * n_local * n_in + 8 * stride
    n_sets = max(1, min(8, -(-300_000_000 // max(set_bytes, 1)))) if n_local else 1
    sets = []
    for i in range(n_sets):
        if n_local:
            x = torch.stack([ctx.device_noise(n, 100 + s + 100000 * i, sigma, n_in) for s in mine])
            dirs = [lissajous(s + 4096 * i, fs)(times) for s in mine]
            elev = torch.from_numpy(np.ascontiguousarray(np.stack([d[0] for d in dirs]))).to(ctx.dev)
            azim = torch.from_numpy(np.ascontiguousarray(np.stack([d[1] for d in dirs]))).to(ctx.dev)
            job = ah.DeviceRender(torch, bdev, x, n_in, CHUNK, SUB, elev.reshape(-1), azim.reshape(-1), cabi.AZ_F64, mix_mode, args.variant)
        else:
            job = None
        out = torch.zeros((2, stride) if mix_mode else (max(n_local, 1), 2, stride), dtype=torch.float32, device=ctx.dev)
        sets.append((job, out))
    st = ctx.stream.cuda_stream

    def reduce_segment(out, pa, pb, works):
        for ear in range(2):
            piece = out[ear, pa:pb]
            if collective == 'all_reduce':
                works.append(dist.all_reduce(piece, async_op=True))
            elif collective == 'reduce':
                works.append(dist.reduce(piece, dst=0, async_op=True))

    def step(i, replicate=None, pipeline=None):
        job, out = sets[i % n_sets]
        works = []
        pipe = peer is not None and (pipelined if pipeline is None else pipeline)
        if job is not None:
            job.plan(st)
        if peer is not None and not pipe:
            peer.begin(st)                                # sharded results: the owners have finished with the previous step
        for k, (pa, pb) in enumerate(segs):
            if job is not None:
                # pipelined: the receive buffers of this step's parity (free once the step three back has been summed
                # everywhere); the last render of the step signals the owners from inside the kernel
                route = peer.submit_route(st, signal=k == len(segs) - 1) if pipe else peer.route if peer is not None else None
                job.render(st, pa, pb, out.data_ptr() + 4 * pa, stride, normalise=not mix_mode, route=route)
            if collective in ('all_reduce', 'reduce'):
                reduce_segment(out, pa, pb, works)        # NCCL stream: ordered after this render, beside the next one
        if pipe:
            # the exchange (wait for all writers, rank-order sum of this rank's slice, wait for all owners) runs on a
            # side stream beside the plan and render of the next step
            peer.submit(st, replicate=(not sharded) if replicate is None else replicate, rendered=job is not None)
        elif peer is not None:
            if job is None:
                peer.zero_my_blocks()
            # signal, rank-order sum of this rank's slice; replicated: stored on every rank, wait for all slices
            peer.finish(st, replicate=(not sharded) if replicate is None else replicate)
        for w in works:
            w.wait()

    def drain():
        """End of a run of steps: the launching stream waits for the exchanges still in flight on the side stream."""
        if peer is not None:
            peer.flush(st)

    def mix_of(i):
        """The global mix step i left behind (assembled from the ranks' slices when it was left sharded)."""
        if peer is None:
            return sets[i % n_sets][1]
        peer.flush(st)
        return peer.gather() if sharded else peer.result

    def step_late_collective(i):
        """The round-1 arrangement, for comparison: one collective after the last render."""
        job, out = sets[i % n_sets]
        if job is not None:
            job.plan(st)
            for pa, pb in segs:
                job.render(st, pa, pb, out.data_ptr() + 4 * pa, stride)
        if collective == 'reduce':
            dist.reduce(out, dst=0)
        elif collective != 'none':
            dist.all_reduce(out)

    def render_only(i):
        job, out = sets[i % n_sets]
        if job is not None:
            for pa, pb in segs:
                job.render(st, pa, pb, out.data_ptr() + 4 * pa, stride)

    warmup = max(args.warmup, 3)
    with ClockSampler(ctx.local_rank) as clocks:
        ms_step, per_step = ctx.timed(step, args.steps, warmup, after=drain)
        if len(clocks.samples) < 20:                     # clock evidence only; not part of any reported time
            t_end = time.time() + 0.5
            while time.time() < t_end:
                for i in range(4):
                    step(i)
                torch.cuda.synchronize()
    value = n_src * n_out / (ms_step * 1e-3)
    # nothing was skipped: no trajectory error, and no source peak above 1 (the rare second pass of
    # apply_hrtf.py:462-464 for mixes would have been due otherwise)
    for job, _ in sets:
        if job is not None:
            small = job.small.cpu().numpy()
            assert cabi.decode_status(small)[0] == 0
            assert not mix_mode or float(small[2:].view(np.float32).max()) <= 1.0
    per_rank = ctx.gather_objects({'rank': rank, 'mean_ms': float(np.mean(per_step)), 'min_ms': float(np.min(per_step)),
                                   'median_ms': float(np.median(per_step)), 'max_ms': float(np.max(per_step)), 'sources': n_local})
    # plan_build + renders, + the peer kernels: pipelined = the reduce (its waits are stream memory operations, or two
    # wait kernels where the driver has none); inside the step = wait + reduce
    peer_kernels = 0 if peer is None else (1 if peer._stream_wait_ok else 3) if pipelined else 2
    launches_per_step = ((1 + len(segs)) if n_local else 0) + peer_kernels
    if not mix_mode:
        launches_per_step += n_local

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak' if weak else 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(name), 'sources': n_src, 'sources_per_gpu': [p['sources'] for p in per_rank],
                   'parallelism': ('one source per rank, no data-path collective' if weak else
                                   'sources sharded round-robin over %d rank(s); per-rank mix in the render kernel; sum of the (2, N_out) fp32 '
                                   'mixes: %s' % (world, {'peer': 'fused with the render over NVLink peer memory - every finished tile is stored '
                                                          'into its owner rank\'s receive buffer, then a rank-order reduce kernel per owner that stores the sum '
                                                          'on EVERY rank (all-reduce semantics; csrc/peer.cu)',
                                                          'peer_sharded': 'fused with the render over NVLink peer memory - every finished tile is stored into its '
                                                          'owner rank\'s receive buffer, then a rank-order reduce kernel per owner; the summed mix is left SHARDED BY '
                                                          'TIME over the ranks (reduce-scatter semantics; csrc/peer.cu) - collective.ms_per_step_replicated is the '
                                                          'all-reduce form',
                                                          'peer_pipelined': 'fused with the render over NVLink peer memory - every finished tile is stored into its '
                                                          'owner rank\'s receive buffer (three sets, by step number) and the render\'s last CTA signals the owners; a '
                                                          'rank-order reduce kernel per owner on a side stream, BESIDE the plan and render of the next step; the summed '
                                                          'mix is left SHARDED BY TIME over the ranks (reduce-scatter semantics; csrc/peer.cu).  The timed region ends '
                                                          'when the last step\'s exchange is complete on every rank.  collective.ms_per_step_exchange_not_pipelined / '
                                                          '_replicated: the exchange inside its own step / all-reduce form',
                                                          'none': 'none (one rank)'}.get(collective, 'NCCL %s per time segment (%d segments)' % (collective, len(segs))))),
                   'value_counts': 'source-sample-pairs/s: every source contributes N_out = %d output pairs per step' % n_out,
                   'l2_policy': 'steps rotate over %d input/output set(s) of %.0f MB per rank (L2: 126 MB)' % (n_sets, set_bytes / 1e6),
                   'kernels_per_step': 'memset, plan_build, %d x render (filter rows synthesised by producer warps in-kernel)%s' % (
                       len(segs), (', peer_reduce on a side stream (waits for the peers: cuStreamWaitValue32)' if pipelined and peer._stream_wait_ok else
                                   ', peer_reduce (+ peer_wait)') if peer is not None else
                       ', 2 x %d NCCL %s' % (len(segs), collective) if collective != 'none' else '')},
        'clocks': clocks.summary(), 'gpu_launches': launches_per_step * args.steps,
        'per_rank_step_ms': {'min_of_means': min(p['mean_ms'] for p in per_rank), 'median_of_means': float(np.median([p['mean_ms'] for p in per_rank])),
                             'max_of_means': max(p['mean_ms'] for p in per_rank), 'ranks': per_rank},
    }

    # ---- comparison: the same step with one collective after the last render ------------------------
    if collective != 'none':
        ms_late, _ = ctx.timed(step_late_collective, max(5, args.steps // 2), 3)
        ms_nocomm, _ = ctx.timed(lambda i: (sets[i % n_sets][0].plan(st) if sets[i % n_sets][0] else None, render_only(i)), max(5, args.steps // 2), 3)
        ms_repl = None
        if peer is not None:
            ms_repl, _ = ctx.timed(lambda i: step(i, replicate=True), max(5, args.steps // 2), 3, after=drain)
            ms_serial = None
            if pipelined:
                ms_serial, _ = ctx.timed(lambda i: step(i, pipeline=False), max(5, args.steps // 2), 3, after=drain)
            if sharded:
                step(0)                                  # leave the buffers in the default (sharded) state
                drain()
        line['collective'] = {'op': collective, 'bytes_per_step': int(8 * n_out), 'ms_per_step': ms_step, 'ms_per_step_replicated': ms_repl,
                              'ms_per_step_exchange_not_pipelined': ms_serial,
                              'ms_per_step_with_one_nccl_all_reduce_at_the_end': ms_late, 'ms_per_step_without_exchange': ms_nocomm,
                              'exposed_ms': ms_step - ms_nocomm}

    # ---- parity -------------------------------------------------------------------------------------
    if mix_mode:
        line['parity'] = mix_parity(ctx, bank, sets[0], name, mine, segs, step, mix_of)

    # ---- roofline of the dominant kernel (rank 0) -----------------------------------------------------
    if rank == 0 and n_local:
        ms_render, _ = ctx.timed(render_only, max(5, args.steps // 2), 3, collective=False)
        hbm_peak, peak_src = ctx.hbm_peak()
        algo = (4.0 * n_local + 8.0) * n_out if mix_mode else 12.0 * n_out * n_local
        useful_fma = 2.0 * k * n_in * n_local
        fma = ctx.fma_probe()
        traffic, traffic_src = traffic_of(mix_mode)
        line['roofline'] = {
            'bound': 'hbm', 'kernel': 'bas_render_tiled_kernel<MIX, FUSED> (%d launches per step, one per time segment)' % len(segs),
            'achieved': algo / (ms_render * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': algo / (ms_render * 1e-3) / 1e9 / hbm_peak,
            'traffic': traffic, 'traffic_source': traffic_src, 'ms_per_step_all_launches': ms_render, 'peak_source': peak_src,
            'algorithmic_bytes_per_step': algo,
            'note': 'algorithmic bytes of a mixing launch: 4 B per source sample read + 8 B per mixed output pair written; the kernel is bound '
                    'by the FP32 FMA pipe (2 K FMAs per source sample against 4 bytes), so its HBM fraction falls as sources are mixed '
                    'on-chip; single_source.roofline is the figure comparable with round 1',
            'fp32_pipe': {'achieved_tfma_s': useful_fma / (ms_render * 1e-3) / 1e12, 'peak_measured_tfma_s': max(fma['fma_f32'], fma['fma_f32x2']),
                          'peak_nominal_tfma_s': NOMINAL_TFMA, 'frac_of_nominal': useful_fma / (ms_render * 1e-3) / 1e12 / NOMINAL_TFMA, 'probe': fma},
        }

    # ---- e2e through the public API with host arrays ---------------------------------------------------
    line['e2e'] = mix_e2e(ctx, bank, name, mine, n, n_out, n_pts, sigma, fs, mix_mode)

    if rank == 0:
        if name != 'single' and not args.no_single:
            line['single_source'] = single_source_block(ctx, make_bank(8, 256), 44100, 60, 0.05, min(args.steps, 50))
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline(ctx, bank, name, sets[0], mine, fs, sigma, n_out)
        print(json.dumps(line))


def mix_parity(ctx, bank, first_set, name, mine, segs, step, mix):
    """The N-GPU mix of step 0's inputs against (a) rank 0 rendering ALL sources alone in one launch over a
    window, (b) the sum of the sources rendered one by one through the non-mixing kernel (each rank its own,
    summed in float64 across ranks)."""
    torch, bas, dist = ctx.torch, ctx.bas, ctx.dist
    idx, n_src, secs, fs, keep, ups, sigma = CONFIGS[name]
    n = secs * fs
    k, n_in, n_out = bas.render_geometry(n, CHUNK, SUB, bank)
    times = np.arange(0, n_in + 1, CHUNK, dtype=np.int64)
    job, _ = first_set
    step(0)                                              # sets[0] again
    ctx.barrier()
    out = mix(0)                                         # the summed mix of set 0
    windows = [(0, 16384), (n_out // 2 // 8192 * 8192 - 5000, n_out // 2 // 8192 * 8192 + 11384), (n_out - 16384, n_out)]
    res = {'windows': windows, 'tolerance_rel_l2': 1e-6}
    # (b) per-source renders, non-mixing kernel
    worst_b = 0.0
    for p0, p1 in windows:
        if job is not None:
            elev = job.elev_d.reshape(len(mine), -1)
            azim = job.azim_d.reshape(len(mine), -1)
            per = bas.render_sources(job.x, CHUNK, SUB, (elev, azim, ctx.cabi.AZ_F64), bank, mix=False, normalise=False,
                                     return_device=True, time_range=(p0, p1))
            local = per.double().sum(dim=0)
        else:
            local = torch.zeros((2, p1 - p0), dtype=torch.float64, device=ctx.dev)
        if dist is not None:
            dist.all_reduce(local)
        got = out[:, p0:p1].double()
        if ctx.args.collective == 'reduce' and dist is not None and ctx.rank != 0:
            continue
        worst_b = max(worst_b, float((got - local).norm() / local.norm()))
    res['mix_vs_sum_of_single_source_renders_rel_l2'] = ctx.max_over_ranks(worst_b)
    # (a) rank 0 alone, all sources in one mixing launch
    worst_a = 0.0
    if ctx.rank == 0:
        x_all = torch.stack([ctx.device_noise(n, 100 + s, sigma, n_in) for s in range(n_src)])
        dirs = [lissajous(s, fs)(times) for s in range(n_src)]
        elev = torch.from_numpy(np.ascontiguousarray(np.stack([d[0] for d in dirs]))).to(ctx.dev)
        azim = torch.from_numpy(np.ascontiguousarray(np.stack([d[1] for d in dirs]))).to(ctx.dev)
        for p0, p1 in windows:
            alone = bas.render_sources(x_all, CHUNK, SUB, (elev, azim, ctx.cabi.AZ_F64), bank, mix=True, normalise=False,
                                       return_device=True, time_range=(p0, p1)).double()
            got = out[:, p0:p1].double()
            worst_a = max(worst_a, float((got - alone).norm() / alone.norm()))
        del x_all
    res['n_gpu_mix_vs_rank0_alone_rel_l2'] = ctx.max_over_ranks(worst_a)
    res['ok'] = bool(res['n_gpu_mix_vs_rank0_alone_rel_l2'] <= 1e-6 and res['mix_vs_sum_of_single_source_renders_rel_l2'] <= 1e-6)
    return res


def mix_e2e(ctx, bank, name, mine, n, n_out, n_pts, sigma, fs, mix_mode):
    """The job through the public API with host arrays, on every rank (its own sources, its own PCIe link):
    pageable ndarrays and plain lambdas in; the mix as a host array on rank 0 out."""
    torch, bas, dist = ctx.torch, ctx.bas, ctx.dist
    from binaural_audio_synthesis_b200 import distributed
    n_local = len(mine)
    x_host = np.stack([noise_host(n, 100 + s, sigma) for s in mine]) if n_local else np.zeros((0, n), dtype=np.float32)
    trajs = [lissajous(s, fs) for s in mine]
    pinned_out = torch.empty((2, n_out), dtype=torch.float32, pin_memory=True) if ctx.rank == 0 else None

    def call(x, fns):
        if not mix_mode:
            return bas.make_signal_move_2d(x[0], CHUNK, SUB, fns[0], bank)
        if dist is None:
            return bas.render_sources(x, CHUNK, SUB, fns, bank, mix=True)
        mix = distributed.render_mix_by_source(x, CHUNK, SUB, fns, bank, dst=0)
        if mix is not None:
            pinned_out.copy_(mix, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return pinned_out.numpy()
        return None

    def measure(x, fns, reps):
        for _ in range(2):
            call(x, fns)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            y = call(x, fns)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        return ctx.max_over_ranks((time.perf_counter() - t0) / reps), y

    reps = max(3, min(ctx.args.steps, 10 if mix_mode else 30))
    dt, y = measure(x_host, trajs, reps)
    n_src_total = CONFIGS[name][1] if mix_mode else ctx.world
    e2e = {'value': n_src_total * n_out / dt, 'unit': UNIT, 'ms_per_step': 1e3 * dt, 'steps': reps, 'n_gpus': ctx.world,
           'h2d_bytes_per_step': int(n_src_total * (4 * n + 16 * n_pts)), 'd2h_bytes_per_step': int(8 * n_out * (1 if mix_mode else ctx.world)),
           'call': ('distributed.render_mix_by_source(pageable float32 ndarray (n_local, N), 512, 32, plain lambdas, bank, dst=0) on every rank + '
                    'download of the mix on rank 0' if dist is not None and mix_mode else
                    'render_sources(pageable float32 ndarray (%d, N), 512, 32, plain lambdas, bank, mix=True) -> host (2, N_out)' % n_local if mix_mode else
                    'make_signal_move_2d(pageable float32 ndarray, 512, 32, plain lambda, bank) -> host ndarray, one call per rank') +
                   '; wall clock, max over ranks; arrays are page-locked in place on their first call (before the timed region) and re-used',
           'cpu_affinity': ctx.cpu_note}
    if mix_mode:
        # the opt-in form: pinned inputs, trajectories declared vectorised
        x_pin = torch.from_numpy(x_host).pin_memory().numpy() if n_local else x_host
        fast = []
        for f in trajs:
            g = (lambda t, f=f: f(t))
            g.vectorized = True
            fast.append(g)
        dt_fast, _ = measure(x_pin, fast, reps)
        e2e['e2e_fast'] = {'value': n_src_total * n_out / dt_fast, 'ms_per_step': 1e3 * dt_fast,
                           'call': 'the same with inputs already pinned and trajectories declared vectorised'}
    return e2e


def cpu_baseline(ctx, bank, name, first_set, mine, fs, sigma, n_out):
    """The reference's own make_signal_move_2d, one host core, on a bounded sample of the job: the first
    stretch of EVERY source, summed like the mix - which also checks the GPU mix against the reference."""
    torch = ctx.torch
    render, kind = cpu_renderer()
    idx, n_src, secs, fs, keep, ups, _ = CONFIGS[name]
    job, out = first_set
    budget_samples = 3_000_000 if kind == 'reference' else 12_000_000      # about 15 s of CPU work
    stretch = max(CHUNK * 4, min(secs * fs, budget_samples // n_src) // CHUNK * CHUNK)
    x_dev = job.x[:, :stretch].cpu().numpy()
    obank = _bank_object((ups, bank.diffs_left, bank.diffs_right, bank.irs_left, bank.irs_right))
    t0 = time.perf_counter()
    acc = np.zeros((stretch + keep - 1, 2), dtype=np.float64)
    for i, s in enumerate(mine):
        acc += render(x_dev[i], CHUNK, SUB, lissajous(s, fs), obank)
    dt = time.perf_counter() - t0
    valid = stretch - 1                                            # outputs below `stretch` depend on inputs below it only
    if name == 'single':
        got = out[0, :, :valid].cpu().numpy().T.astype(np.float64)
    else:
        got = out[:, :valid].cpu().numpy().T.astype(np.float64)
    want = acc[:valid]
    rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    peak = float(np.abs(got - want).max() / np.abs(want).max())
    return {'value': n_src * stretch / dt, 'unit': UNIT, 'cores': 1, 'kind': kind,
            'sample': 'the first %d samples (%.2f s) of each of the %d sources, one process (%s)' % (
                stretch, stretch / fs, n_src, 'unmodified reference, oracle/_ref' if kind == 'reference' else 'numpy port of oracle/'),
            'gpu_vs_cpu_rel_l2': rel, 'gpu_vs_cpu_max_abs_over_peak': peak, 'tolerance': 1e-5}


def run_hour(ctx):
    """configs[3]: one 1-hour 48 kHz source cut across the ranks by time.  A step: every rank plans and renders the
    output samples of its own segment from its inputs plus the K-1 halo, then one MAX all-reduce of the peak and
    the division (apply_hrtf.py:462-464)."""
    args, torch, bas, cabi, lib, dist = ctx.args, ctx.torch, ctx.bas, ctx.cabi, ctx.lib, ctx.dist
    ah = bas.apply_hrtf
    from binaural_audio_synthesis_b200 import distributed
    idx, _, secs, fs, keep, ups, sigma = CONFIGS['hour']
    rank, world = ctx.rank, ctx.world
    bank = make_bank(ups, keep)
    bdev = ah._device_bank(bank)
    n = secs * fs
    k, n_in, n_out = bas.render_geometry(n, CHUNK, SUB, bank)
    ranges = distributed.time_segments(n_in, CHUNK, k, world)
    p0, p1 = ranges[rank]
    n0, n1 = distributed.segment_inputs(p0, p1, n_in, CHUNK, k)
    block = 1 << 20                                               # the global signal: blocks of 2^20 samples, block b seeded 40000 + b

    def signal_window(a, b):
        parts = []
        for blk in range(a // block, (b - 1) // block + 1):
            parts.append(ctx.device_noise(block, 40000 + blk, sigma))
        x = torch.cat(parts)[a - a // block * block: a - a // block * block + (b - a)]
        if b > n:
            x[n - a:] = 0
        return x.contiguous()

    traj = spiral(fs, secs)
    kind = bas.sphere.az_kind(traj(0)[1])

    def directions(a, b):
        t = np.arange(a, b + 1, CHUNK, dtype=np.int64)
        e, az = traj(t)
        return (torch.from_numpy(np.ascontiguousarray(np.broadcast_to(e, t.shape))).to(ctx.dev),
                torch.from_numpy(np.ascontiguousarray(np.broadcast_to(az, t.shape))).to(ctx.dev))

    x = signal_window(n0, n1)[None, :]
    elev, azim = directions(n0, n1)
    job = ah.DeviceRender(torch, bdev, x, n1 - n0, CHUNK, SUB, elev, azim, kind, False, args.variant)
    q0, q1 = p0 - n0, min(p1, n1 + k - 1) - n0
    stride = (q1 - q0 + 3) // 4 * 4
    out = torch.empty((1, 2, stride), dtype=torch.float32, device=ctx.dev)
    peak = torch.zeros(1, dtype=torch.float32, device=ctx.dev)
    st = ctx.stream.cuda_stream

    def step(i):
        job.plan(st)
        job.render(st, q0, q1, out.data_ptr(), stride)
        peak.copy_(job.peaks[:1])
        if dist is not None:
            dist.all_reduce(peak, op=dist.ReduceOp.MAX)
        cabi.check(lib.bas_normalise(out.data_ptr(), 2 * stride, peak.data_ptr(), st), 'bas_normalise')

    warmup = max(args.warmup, 3)
    with ClockSampler(ctx.local_rank) as clocks:
        ms_step, per_step = ctx.timed(step, args.steps, warmup, after=drain)
    assert cabi.decode_status(job.small.cpu().numpy())[0] == 0
    per_rank = ctx.gather_objects({'rank': rank, 'mean_ms': float(np.mean(per_step)), 'outputs': int(p1 - p0)})
    # parity: every seam against rank 0 rendering a window across it alone; the first window against the oracle
    seam = 4096
    parity = {'seams': [], 'tolerance_rel_l2': 1e-6}
    edges = [ranges[r][0] for r in range(1, world) if ranges[r][1] > ranges[r][0]]
    for c in edges:
        piece = torch.zeros((2, 2 * seam), dtype=torch.float32, device=ctx.dev)
        lo, hi = max(p0, c - seam), min(p1, c + seam)
        if hi > lo:
            piece[:, lo - (c - seam): hi - (c - seam)] = out[0, :, lo - p0: hi - p0]
        if dist is not None:
            dist.all_reduce(piece)                                 # disjoint halves from the two neighbours
        if rank == 0:
            a = (c - seam - (k - 1)) // CHUNK * CHUNK
            b = (c + seam + CHUNK - 1) // CHUNK * CHUNK
            xw = signal_window(a, b)[None, :]
            ew, aw = directions(a, b)
            alone = bas.render_sources(xw, CHUNK, SUB, (ew[None], aw[None], kind), bank, normalise=False, return_device=True,
                                       time_range=(c - seam - a, c + seam - a))[0]
            parity['seams'].append(float((piece.double() - alone.double()).norm() / alone.double().norm()))
    parity['ok'] = all(v <= 1e-6 for v in parity['seams'])
    line = {
        'metric': METRIC, 'value': n_out / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name('hour'), 'parallelism': 'time segments at multiples of chunksize, K-1 input halo, no tail exchange; '
                   'MAX all-reduce of the peak (1 float); output left sharded by time', 'outputs_per_rank': [p['outputs'] for p in per_rank],
                   'l2_policy': 'inputs larger than L2 (%.0f MB per rank)' % (4 * (n1 - n0) / 1e6)},
        'clocks': clocks.summary(), 'gpu_launches': 3 * args.steps, 'parity': parity,
        'per_rank_step_ms': per_rank,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            render, kind_cpu = cpu_renderer()
            stretch = 65536
            xs = x[0, :stretch].cpu().numpy()
            t0 = time.perf_counter()
            want = render(xs, CHUNK, SUB, traj, _bank_object((ups, bank.diffs_left, bank.diffs_right, bank.irs_left, bank.irs_right)))
            dt = time.perf_counter() - t0
            got = out[0, :, :stretch - 1].cpu().numpy().T.astype(np.float64)
            line['cpu_baseline'] = {'value': stretch / dt, 'unit': UNIT, 'cores': 1, 'kind': kind_cpu, 'sample': 'first %d samples' % stretch,
                                    'gpu_vs_cpu_rel_l2': float(np.linalg.norm(got - want[:stretch - 1]) / np.linalg.norm(want[:stretch - 1]))}
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='mix64', help='mix64 (default) | single | hour | stress1024, or SURVEY.md config number 2..5')
    ap.add_argument('--variant', type=lambda v: int(v, 0), default=0, help='bas_render variant (tuning)')
    ap.add_argument('--segments', type=int, default=0, help='time segments of the mix (0: distributed.MIX_SEGMENTS)')
    ap.add_argument('--collective', default='peer_pipelined', choices=['peer_pipelined', 'peer_sharded', 'peer', 'all_reduce', 'reduce'],
                    help='sum of the per-rank mixes: fused with the render over peer memory, result left sharded by time, the exchange of a '
                         'step beside the render of the next (default) or inside its own step (peer_sharded), or replicated on every rank '
                         '(peer), or NCCL after the render')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-single', action='store_true', help='skip the single_source block')
    args = ap.parse_args()
    args.config = ALIASES.get(args.config, args.config)
    if args.config not in CONFIGS:
        ap.error('unknown config %r' % args.config)
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    ctx = Ctx(args, rank, local_rank, world)
    try:
        if args.config == 'hour':
            run_hour(ctx)
        else:
            run_mix(ctx, args.config)
    finally:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()


if __name__ == '__main__':
    main()
