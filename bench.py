#!/usr/bin/env python
"""Benchmark of the moving-source binaural render path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] - one 60 s mono 44.1 kHz pink-noise source on a
Lissajous azimuth+elevation trajectory, N=8 upsampled synthetic bank, samples_to_keep=256,
chunksize=512, subchunksize=32.  A step = one full render of that source:
    plan_build (5169 directions) -> ir_synth -> render (FIR + crossfade) -> normalise.
At N GPUs every rank renders its own source of that shape (weak scaling, sources are independent,
no collective on the data path).

value      output sample-pairs/s over all ranks, inputs resident in HBM, CUDA events, max over ranks.
           Steps rotate over buffer sets whose total size exceeds L2 (config.l2_policy).
e2e        the same metric through the public call make_signal_move_2d(host ndarray) -> host
           ndarray, host<->device copies inside the timed region.
roofline   the render kernel alone: algorithmic HBM bytes (12 B per output pair: 4 B in, 8 B out)
           over its CUDA-event duration, against MEASURED_PEAKS.json's hbm_gbs; plus, because the
           path is bound by the FP32 pipe, the same launch as FMA/s against an FMA
           peak measured here with bas_probe_fma.
cpu_baseline   the numpy oracle port of the reference (oracle/binaural_oracle.py), 1 core, on the
           first seconds of the same workload.
--impl reference   the oracle port on all host cores (one process per time segment).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 44100
SECONDS = 60
CHUNK, SUB, KEEP, UPS = 512, 32, 256, 8
METRIC = 'binaural output sample-pairs/s'
UNIT = 'sample-pairs/s'


def workload_name():
    return ('configs[1]: 60 s mono 44.1 kHz pink noise, Lissajous az+el trajectory, N=%d bank, samples_to_keep=%d, '
            'chunksize=%d, subchunksize=%d' % (UPS, KEEP, CHUNK, SUB))


def pink_noise(n, seed):
    """1/f-shaped Gaussian noise, sigma = 0.05 (peak stays below 1: apply_hrtf.py:462 inert)."""
    rng = np.random.default_rng(seed)
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.arange(spec.size, dtype=np.float64)
    f[0] = 1.0
    x = np.fft.irfft(spec / np.sqrt(f), n)
    return (0.05 * x / x.std()).astype(np.float32)


def lissajous(seed=0):
    rng = np.random.default_rng(1000 + seed)
    p1, p2 = (0.3, 1.0) if seed == 0 else (rng.uniform(0, 6), rng.uniform(0, 6))
    k = 2 * np.pi / (4 * FS)

    def fn(t):
        return (np.deg2rad(22.5 + 67.5 * np.sin(3 * k * t + p1)), (5 * k * t + p2) % (2 * np.pi))
    fn.vectorized = True
    return fn


def make_bank(bas):
    f = bas.bank_synth.build_bank(UPS, seed=0)

    class bank:
        upsampling = UPS
        diffs_left, diffs_right = f['diffs_left'], f['diffs_right']
        irs_left, irs_right = f['irs_left'][:, :KEEP * UPS], f['irs_right'][:, :KEEP * UPS]
    return bank


def pin_to_device_cpus(index):
    """Run this rank on the CPU cores NVML names as local to GPU `index` (same NUMA node / PCIe root):
    the pinned host buffers of the e2e leg are then allocated next to the GPU they feed."""
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
# reference arm: the numpy port of the reference on all host cores
# ---------------------------------------------------------------------------------------------------
def _cpu_segment(args):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import binaural_oracle as oracle
    x, t0, bank_fields, seed = args
    bank = oracle.Bank(*bank_fields)
    traj = lissajous(seed)
    return oracle.make_signal_move_2d(x, CHUNK, SUB, lambda t: traj(np.float64(t + t0)), bank).shape[0]


def run_reference(args, rank, world):
    """bench.py --impl reference: the reference's algorithm (numpy port, oracle/) on every host core.
    A step renders `cores` independent 1.5 s stretches of the workload signal, one per process."""
    if rank != 0:
        return
    import multiprocessing as mp
    import binaural_audio_synthesis_b200.bank_synth as bank_synth
    f = bank_synth.build_bank(UPS, seed=0)
    fields = (UPS, f['diffs_left'], f['diffs_right'], f['irs_left'][:, :KEEP * UPS], f['irs_right'][:, :KEEP * UPS])
    cores = os.cpu_count() or 1
    x = pink_noise(SECONDS * FS, 2)
    seg = int(1.5 * FS) // CHUNK * CHUNK
    starts = [(i * seg) % (x.size - seg) // CHUNK * CHUNK for i in range(cores)]
    jobs = [(x[s:s + seg], s, fields, 0) for s in starts]
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(_cpu_segment, jobs)
        t0 = time.perf_counter()
        pairs = 0
        for _ in range(args.steps):
            pairs += sum(pool.map(_cpu_segment, jobs))
        dt = time.perf_counter() - t0
    value = pairs / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(), 'parallelism': 'host processes, one per time segment'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d stretches of %.2f s of the workload signal per step, one per process' % (cores, seg / FS)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import binaural_audio_synthesis_b200 as bas
    from binaural_audio_synthesis_b200 import _cabi
    lib = _cabi.lib
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    cpu_note = pin_to_device_cpus(local_rank) if world > 1 and not os.environ.get('BAS_NO_CPU_AFFINITY') else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    bas.apply_hrtf.PROGRESS = False
    bank = make_bank(bas)
    bdev = bas.apply_hrtf._device_bank(bank)
    n = SECONDS * FS
    k, n_in, n_out = bas.render_geometry(n, CHUNK, SUB, bank)
    n_pts = n_in // CHUNK + 1
    stream = torch.cuda.current_stream().cuda_stream

    # buffer sets: rotating over more bytes than L2 holds, so every step streams from HBM
    n_sets = 8
    pitch = lib.bas_filter_row_pitch(k)
    set_bytes = 4 * (n_in + 2 * n_out + n_pts * 2 * pitch)
    times = np.arange(0, n_in + 1, CHUNK, dtype=np.int64)
    sets = []
    for i in range(n_sets):
        x = torch.zeros(n_in, dtype=torch.float32, device=dev)
        x[:n] = torch.from_numpy(pink_noise(n, 2 + 100 * rank + i)).to(dev)
        elev, azim = lissajous(0 if (rank == 0 and i == 0) else 1 + 100 * rank + i)(times)
        sets.append(dict(
            x=x, elev=torch.from_numpy(np.ascontiguousarray(elev)).to(dev), azim=torch.from_numpy(np.ascontiguousarray(azim)).to(dev),
            terms=torch.empty(n_pts * 256, dtype=torch.uint8, device=dev), status=torch.zeros(2, dtype=torch.int32, device=dev),
            filt=torch.empty((n_pts, pitch, 2), dtype=torch.float32, device=dev),
            out=torch.empty((2, (n_out + 4) // 4 * 4), dtype=torch.float32, device=dev), peak=torch.zeros(1, dtype=torch.float32, device=dev)))
    out_stride = (n_out + 4) // 4 * 4
    # sources in flight: consecutive steps (independent sources) alternate over this many streams, the
    # way a server keeps several make_signal_move_2d calls going on one GPU.  plan_build (latency
    # bound), ir_synth (L2 bound) and render (FP32 pipe bound) of neighbouring steps then overlap, and
    # the next step's CTAs fill the SMs the last wave of a render leaves idle.
    n_flight = max(1, args.in_flight)
    streams = [torch.cuda.current_stream()] + [torch.cuda.Stream() for _ in range(n_flight - 1)]
    workspaces = [torch.empty(int(lib.bas_render_workspace_bytes()), dtype=torch.uint8, device=dev) for _ in streams]

    def step_render(s, lane=0):
        st = streams[lane].cuda_stream
        _cabi.check(lib.bas_render(s['x'].data_ptr(), n_in, n_in, 1, n_in, CHUNK, SUB, k, s['filt'].data_ptr(), None,
                                   0, n_out, s['out'].data_ptr(), out_stride, 0, s['peak'].data_ptr(), args.variant,
                                   workspaces[lane].data_ptr(), workspaces[lane].numel(), st), 'bas_render')

    def step(s, lane=0):
        st = streams[lane].cuda_stream
        _cabi.check(lib.bas_plan_build(bdev.diffs[0].data_ptr(), bdev.diffs[1].data_ptr(), UPS, k * UPS, s['elev'].data_ptr(),
                                       s['azim'].data_ptr(), None, _cabi.AZ_F64, n_pts, s['terms'].data_ptr(), None,
                                       s['status'].data_ptr(), st), 'bas_plan_build')
        _cabi.check(lib.bas_ir_synth(bdev.bank_pp.data_ptr(), UPS, k, s['terms'].data_ptr(), n_pts, _cabi.IR_ROWS,
                                     s['filt'].data_ptr(), k, st), 'bas_ir_synth')
        _cabi.check(lib.bas_memset(s['peak'].data_ptr(), 0, 4, st), 'bas_memset')
        step_render(s, lane)
        _cabi.check(lib.bas_normalise(s['out'].data_ptr(), 2 * out_stride, s['peak'].data_ptr(), st), 'bas_normalise')
    launches_per_step = 7        # plan (2 status memsets + kernel), ir_synth, peak memset, render, normalise

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, collective=True, lanes=1):
        """ms per step of fn over `steps` steps after `warmup`: CUDA events on the launching stream,
        bracketed by a barrier + synchronize, max over ranks.  lanes > 1: step i runs on stream
        i % lanes; the closing event is recorded after every stream has joined the first.
        collective=False: this rank only (the per-kernel timings rank 0 takes for the roofline)."""
        sync = barrier if collective else torch.cuda.synchronize
        for i in range(warmup):
            fn(sets[i % n_sets], i % lanes) if lanes > 1 else fn(sets[i % n_sets])
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for lane in range(1, lanes):
            streams[lane].wait_event(e0)
        for i in range(steps):
            fn(sets[(warmup + i) % n_sets], i % lanes) if lanes > 1 else fn(sets[(warmup + i) % n_sets])
        for lane in range(1, lanes):
            ev = torch.cuda.Event()
            ev.record(streams[lane])
            streams[0].wait_event(ev)
        e1.record(streams[0])
        sync()
        ms = e0.elapsed_time(e1)
        if collective and dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    warmup = max(args.warmup, 3)
    with ClockSampler(local_rank) as clocks:
        ms_step = timed(step, args.steps, warmup, lanes=n_flight)
        ms_serial = timed(step, args.steps, warmup) if n_flight > 1 else ms_step
        # keep the sampler running over a longer stretch of the same work if the timed region was
        # too short for NVML to see it (clock evidence only; not part of any reported time)
        if len(clocks.samples) < 20:
            t_end = time.time() + 0.5
            while time.time() < t_end:
                for i in range(20):
                    step(sets[i % n_sets], i % n_flight)
                torch.cuda.synchronize()
    assert int(sets[0]['status'].cpu()[0]) == 0
    value = world * n_out / (ms_step * 1e-3)

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_name(), 'sources_per_gpu': 1, 'sources_in_flight': n_flight,
                   'in_flight_note': 'consecutive steps (independent sources) alternate over %d CUDA streams; one source at a time: '
                                     '%.4f ms per step' % (n_flight, ms_serial),
                   'parallelism': 'source-sharded x%d, no data-path collective' % world,
                   'l2_policy': 'steps rotate over %d buffer sets (%.0f MB > 126 MB L2)' % (n_sets, n_sets * set_bytes / 1e6),
                   'kernels_per_step': 'plan_build, ir_synth, render, normalise'},
        'clocks': clocks.summary(), 'gpu_launches': launches_per_step * args.steps,
    }

    if rank == 0:
        # ---- roofline of the dominant kernel (render) -------------------------------------------
        ms_render = timed(step_render, max(args.steps, 20), 3, collective=False)
        ms_synth = timed(lambda s: _cabi.check(lib.bas_ir_synth(bdev.bank_pp.data_ptr(), UPS, k, s['terms'].data_ptr(), n_pts,
                                                              _cabi.IR_ROWS, s['filt'].data_ptr(), k, stream), 'bas_ir_synth'),
                         max(args.steps, 20), 3, collective=False)
        ms_plan = timed(lambda s: _cabi.check(lib.bas_plan_build(bdev.diffs[0].data_ptr(), bdev.diffs[1].data_ptr(), UPS, k * UPS,
                                                              s['elev'].data_ptr(), s['azim'].data_ptr(), None, _cabi.AZ_F64, n_pts,
                                                              s['terms'].data_ptr(), None, s['status'].data_ptr(), stream), 'plan'),
                        max(args.steps, 20), 3, collective=False)
        algo_bytes = 12.0 * n_out
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            hbm_peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        achieved = algo_bytes / (ms_render * 1e-3) / 1e9
        # DRAM bytes of one render launch from the committed ncu --set full capture of this command
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            for name, kv in tj['kernels'].items():
                if name.startswith('bas_render_tiled_kernel'):
                    traffic = kv['dram_bytes_read'] + kv['dram_bytes_write']
                    traffic_src = 'profiles/r1_traffic.json (%s): %s; the 21 MB of output were still in the 126 MB L2 when the kernel ended' % (tj['source'], name)
        # FP32 pipe: measured FMA peak on this GPU, same session
        sink = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)
        fma = {}
        for name, packed in (('fma_f32', 0), ('fma_f32x2', 1)):
            iters = 4096
            ms = timed(lambda s: _cabi.check(lib.bas_probe_fma(packed, 148 * 8, 256, iters, sink.data_ptr(), stream), 'probe'), 5, 2, collective=False)
            fma[name] = 148 * 8 * 256 * iters * 32 / (ms * 1e-3) / 1e12
        # the render kernel's own operand pattern (one tap pair reused along a diagonal, scalar-broadcast x,
        # rotating accumulators) without its loads: what FFMA2 can reach at 12 warps per SM
        ms = timed(lambda s: _cabi.check(lib.bas_probe_fma(4, 148 * 3, 128, 64, sink.data_ptr(), stream), 'probe'), 5, 2, collective=False)
        fma['fma_f32x2_render_pattern'] = 148 * 3 * 128 * 64 * 2048 / (ms * 1e-3) / 1e12
        fma_peak = max(fma['fma_f32'], fma['fma_f32x2'])
        useful_fma = 2.0 * k * n_in                       # 2 ears x K taps per input sample
        line['roofline'] = {
            'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
            'traffic': traffic, 'traffic_source': traffic_src, 'kernel': 'bas_render_tiled_kernel', 'ms_per_launch': ms_render, 'peak_source': peak_src,
            'algorithmic_bytes_per_launch': algo_bytes,
            'fp32_pipe': {'achieved_tfma_s': useful_fma / (ms_render * 1e-3) / 1e12, 'peak_tfma_s': fma_peak,
                          'frac': useful_fma / (ms_render * 1e-3) / 1e12 / fma_peak, 'probe': fma,
                          'note': 'useful FMAs only (2*K per input sample) against the scalar FFMA peak; the kernel issues FFMA2 '
                                  '(leaves issue slots for the loads), whose rate in the kernel\'s operand pattern is probe.fma_f32x2_render_pattern, '
                                  'and 11 % of its FMA-pipe work are tap blends; see DESIGN.md'},
            'ir_synth_ms_per_launch': ms_synth, 'plan_build_ms_per_launch': ms_plan,
        }

    # ---- e2e through the public API with host buffers, on every rank (its own source, its own PCIe
    #      link): pinned host signal in, host array out, copies inside the timed region -------------
    x_pinned = torch.from_numpy(pink_noise(n, 2 + 100 * rank)).pin_memory()      # e2e inputs live in pinned host memory
    x_host = x_pinned.numpy()
    traj = lissajous(0 if rank == 0 else 1 + 100 * rank)
    e2e_steps = max(3, min(args.steps, 50))
    y = None
    for _ in range(max(warmup, 4)):        # results are held like in the timed loop: two pinned result buffers alternate
        y = bas.make_signal_move_2d(x_host, CHUNK, SUB, traj, bank)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y = bas.make_signal_move_2d(x_host, CHUNK, SUB, traj, bank)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    assert y.shape == (n_out, 2)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    line['e2e'] = {'value': world * n_out / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(4 * n + 16 * n_pts),
                   'd2h_bytes_per_step': int(8 * n_out + 12), 'ms_per_step': 1e3 * dt, 'steps': e2e_steps,
                   'call': 'make_signal_move_2d(host float32 ndarray, 512, 32, vectorised trajectory, bank) -> host ndarray, '
                           'one call per rank per step, wall clock, max over ranks',
                   'n_gpus': world, 'cpu_affinity': cpu_note}

    if rank == 0:
        # ---- CPU baseline: numpy port of the reference, one core, first seconds of the workload ----
        if world == 1 and not args.no_cpu:
            from oracle import binaural_oracle as oracle
            obank = oracle.Bank(UPS, bank.diffs_left, bank.diffs_right, bank.irs_left, bank.irs_right)
            sample_s = 4
            xs = x_host[:sample_s * FS]
            t0 = time.perf_counter()
            yo = oracle.make_signal_move_2d(xs, CHUNK, SUB, lambda t: traj(np.float64(t)), obank)   # rank 0: traj = lissajous(0)
            dt = time.perf_counter() - t0
            line['cpu_baseline'] = {'value': yo.shape[0] / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                    'sample': 'first %d s of the workload signal, single process (numpy oracle port)' % sample_s}
            got = y[:xs.size - 1000].astype(np.float64)
            err = np.linalg.norm(got - yo[:xs.size - 1000]) / np.linalg.norm(yo[:xs.size - 1000])
            line['cpu_baseline']['gpu_vs_port_rel_l2'] = float(err)
    if world > 1:
        # the config-3 exchange step, reported beside the data path: SUM-reduce of one (2, N_out) mix
        mix = sets[0]['out']
        for _ in range(3):
            dist.all_reduce(mix)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(mix)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line['mix_allreduce'] = {'ms': float(t), 'bytes': int(mix.numel() * 4),
                                 'note': 'NCCL SUM of the (2, N_out) fp32 mix (config 3 exchange step); not part of value'}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--variant', type=lambda v: int(v, 0), default=0, help='bas_render variant (tuning)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--in-flight', type=int, default=4, help='independent sources (steps) in flight per GPU, one CUDA stream each')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
