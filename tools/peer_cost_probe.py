"""Where does the exchange cost go?  (torchrun, >= 2 GPUs; bench-like mix: [sources per rank] x 60 s.)
Times K back-to-back steps of:  local      plan + render into local memory
                                routed     plan + render routed to the owners (no signal, no reduce)
                                signalled  ... + the in-kernel arrival signal
                                pipelined  ... + the reduce on the side stream (the product path)
                                serial     the exchange inside its own step (begin / finish)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import binaural_audio_synthesis_b200 as bas
import bench

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 64 // world
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ah, cabi = bas.apply_hrtf, bas._cabi
bank = bench.make_bank(16, 256)
bdev = ah._device_bank(bank)
n = 60 * 44100
k, n_in, n_out = bas.render_geometry(n, 512, 32, bank)
stride = (n_out + 3) // 4 * 4
times = np.arange(0, n_in + 1, 512, dtype=np.int64)
dev = torch.device('cuda', local)
g = torch.Generator(device=dev); g.manual_seed(rank)
x = torch.zeros((n_local, n_in), dtype=torch.float32, device=dev)
x[:, :n] = torch.randn((n_local, n), generator=g, device=dev) * 0.01
dirs = [bench.lissajous(rank * n_local + s)(times) for s in range(n_local)]
elev = torch.from_numpy(np.stack([d[0] for d in dirs])).to(dev)
azim = torch.from_numpy(np.stack([d[1] for d in dirs])).to(dev)
job = ah.DeviceRender(torch, bdev, x, n_in, 512, 32, elev.reshape(-1), azim.reshape(-1), cabi.AZ_F64, True, 0)
out = torch.zeros((2, stride), dtype=torch.float32, device=dev)
peer = bas.distributed._peer_mix(n_out, None)
main = torch.cuda.current_stream()
st = main.cuda_stream

def local_step(i):
    job.plan(st); job.render(st, 0, n_out, out.data_ptr(), stride)
def routed(i):
    job.plan(st); job.render(st, 0, n_out, out.data_ptr(), stride, route=peer._routes[i % peer.DEPTH])
def signalled(i):
    r = peer._signal_routes[i % peer.DEPTH]; r.arrive_epoch = 0
    job.plan(st); job.render(st, 0, n_out, out.data_ptr(), stride, route=r)
def pipelined(i):
    job.plan(st); job.render(st, 0, n_out, out.data_ptr(), stride, route=peer.submit_route(st)); peer.submit(st, replicate=False)
def serial(i):
    job.plan(st); peer.begin(st); job.render(st, 0, n_out, out.data_ptr(), stride, route=peer.route); peer.finish(st, replicate=False)
def render_only(i):
    job.render(st, 0, n_out, out.data_ptr(), stride)

def timed(fn, after=None):
    for i in range(4):
        fn(i)
    if after: after()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for i in range(steps):
        fn(i)
    if after: after()
    b.record(main)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    lo = t.clone(); dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return round(float(t), 5), round(float(lo), 5)

def timeline(n_steps=14):
    """Time stamps (one-thread kernels, libbas_probe.so) around the kernels of the pipelined step."""
    import probe_lib
    probe = probe_lib.load()
    slots = torch.zeros((n_steps, 6), dtype=torch.int64, device=dev)
    side = peer._side.cuda_stream
    stamp = lambda stream, i, k: probe.bas_probe_stamp(slots[i, k:].data_ptr(), stream)
    orig = peer._reduce
    for i in range(n_steps):
        stamp(st, i, 0); job.plan(st); stamp(st, i, 1)
        job.render(st, 0, n_out, out.data_ptr(), stride, route=peer.submit_route(st)); stamp(st, i, 2)

        def traced(parity, e, replicate, fold, stream, i=i):
            stamp(side, i, 3); orig(parity, e, replicate, fold, stream); stamp(side, i, 4)
        peer._reduce = traced
        peer.submit(st, replicate=False)
        peer._reduce = orig
        stamp(side, i, 5)
    peer.flush(st)
    torch.cuda.synchronize()
    t = slots.cpu().numpy().astype(np.float64)
    t = (t - t[4, 0]) / 1e3
    return [{'step': i, 'plan_begin': round(t[i, 0], 1), 'plan_end': round(t[i, 1], 1), 'render_end': round(t[i, 2], 1),
             'arrivals_in': round(t[i, 3], 1), 'reduce_end': round(t[i, 4], 1), 'owners_done': round(t[i, 5], 1)} for i in range(4, n_steps)]


sys.path.insert(0, os.path.join(ROOT, 'tools'))
res = {'world': world, 'sources_per_rank': n_local, 'steps': steps}
for rep in range(2):
    for name, fn, after in (('render_only', render_only, None), ('local', local_step, None), ('routed', routed, None), ('signalled', signalled, None),
                            ('pipelined', pipelined, lambda: peer.flush(st)), ('serial', serial, lambda: peer.flush(st))):
        res.setdefault(name + '_ms_max_min', []).append(timed(fn, after))
tl = timeline()
every = [None] * world
dist.all_gather_object(every, tl)
if rank == 0:
    print(json.dumps(res))
    for r, t in enumerate(every):
        for row in t:
            print('rank %d  %s' % (r, json.dumps(row)))
dist.barrier()
dist.destroy_process_group()
