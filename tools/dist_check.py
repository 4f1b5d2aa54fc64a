"""Multi-GPU check of the two shardings (run under torchrun, backend nccl):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
by source: every rank renders and mixes its own sources, NCCL SUM -> compare with one GPU mixing all of them;
by time:   one long source cut across the ranks with an input halo, NCCL MAX of the peak -> compare with
           make_signal_move_2d on one GPU."""
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
bas.apply_hrtf.PROGRESS = False
bank = bench.make_bank(bas)
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
out = {}

# ---- by source (configs 3, 5) ----
n_src, n = 12, 5 * 44100
sig = np.stack([bench.pink_noise(n, 100 + s) for s in range(n_src)])
sig[3] *= 60.0                                     # one source that peaks above 1 and is normalised on its own
trajs = [bench.lissajous(s) for s in range(n_src)]
mine = bas.distributed.shard_sources(n_src, rank, world)
ref = bas.render_sources(sig, 512, 32, trajs, bank, mix=True) if rank == 0 else None
for exchange in ('peer', 'nccl'):
    # twice: the second call re-uses the symmetric buffers and flags of the first (epochs advance)
    for rep in range(2):
        mix = bas.distributed.render_mix_by_source(sig[mine], 512, 32, [trajs[s] for s in mine], bank, exchange=exchange)
    torch.cuda.synchronize()
    every = [None] * world
    dist.all_gather_object(every, float(mix.double().abs().sum()))
    if rank == 0:
        out['by_source_%s_rel_l2_vs_one_gpu' % exchange] = rel(mix.cpu().numpy(), ref.astype(np.float64))
        out['by_source_%s_same_on_every_rank' % exchange] = bool(max(every) == min(every))
        out['by_source_shape'] = list(mix.shape)
# device-resident signals, result only on rank 1 (dst), and a rank count that leaves rank world-1 without sources
few = min(world - 1, 3) if world > 1 else 1
mine_few = bas.distributed.shard_sources(few, rank, world)
xd = torch.from_numpy(sig[mine_few]).cuda() if mine_few else torch.zeros((0, n), dtype=torch.float32, device='cuda')
mix = bas.distributed.render_mix_by_source(xd, 512, 32, [trajs[s] for s in mine_few], bank, dst=world - 1)
if rank == world - 1:
    out_dst = mix.cpu().numpy()
    dist.send(torch.from_numpy(out_dst).cuda(), dst=0) if world > 1 else None
if rank == 0:
    ref_few = bas.render_sources(sig[:few], 512, 32, trajs[:few], bank, mix=True)
    if world > 1:
        got = torch.empty((2, ref_few.shape[1]), dtype=torch.float32, device='cuda')
        dist.recv(got, src=world - 1)
        got = got.cpu().numpy()
    else:
        got = out_dst
    out['by_source_sourceless_rank_rel_l2'] = rel(got, ref_few.astype(np.float64))

# ---- a stream of batches, the exchange of a batch beside the render of the next (PeerMix pipelined) ----
n_b = 4 * 44100
batches, refs = [], []
for b in range(5):
    sb = np.stack([bench.pink_noise(n_b, 900 + 10 * b + s) for s in range(2 * world)])
    tb = [bench.lissajous(7 * b + s) for s in range(2 * world)]
    mine_b = bas.distributed.shard_sources(2 * world, rank, world)
    batches.append((torch.from_numpy(sb[mine_b]).cuda(), [tb[s] for s in mine_b]))
    if rank == 0:
        refs.append(bas.render_sources(sb, 512, 32, tb, bank, mix=True, normalise=False).astype(np.float64))
for replicate in (True, False):
    worst, same = 0.0, True
    for b, got in enumerate(bas.distributed.render_mix_stream(batches, 512, 32, bank, replicate=replicate)):
        got = got.clone()
        if replicate:
            every = [None] * world
            dist.all_gather_object(every, float(got.double().abs().sum()))
            same = same and max(every) == min(every)
            if rank == 0:
                worst = max(worst, rel(got.cpu().numpy(), refs[b]))
        else:
            peer = bas.distributed._peer_mix(got.shape[1], None) if world > 1 else None
            lo, cnt = peer.slice if peer is not None else (0, got.shape[1])
            piece = torch.zeros((2, peer.slice_len if peer is not None else got.shape[1]), dtype=torch.float32, device='cuda')
            piece[:, :cnt] = got[:, lo:lo + cnt]
            pieces = [torch.empty_like(piece) for _ in range(world)]
            dist.all_gather(pieces, piece)
            full = torch.cat(pieces, dim=1)[:, :got.shape[1]]
            if rank == 0:
                worst = max(worst, rel(full.cpu().numpy(), refs[b]))
    if rank == 0:
        out['stream_%s_rel_l2_vs_one_gpu' % ('replicated' if replicate else 'sharded')] = worst
        if replicate:
            out['stream_same_on_every_rank'] = bool(same)
# a failing trajectory on the LAST rank only raises everywhere
bad = [(batches[0][0], [(lambda t: (0.0, float('nan')))] * len(batches[0][1]) if rank == world - 1 else batches[0][1])]
try:
    list(bas.distributed.render_mix_stream(bad, 512, 32, bank))
    raised = False
except (ValueError, AssertionError, bas.BasError):
    raised = True
every = [None] * world
dist.all_gather_object(every, raised)
out['stream_error_raises_on_every_rank'] = bool(all(every))

# ---- by time (config 4) ----
x = 30.0 * bench.pink_noise(20 * 44100 + 123, 7)   # peaks above 1: exercises the global normalisation
traj = bench.lissajous(3)
full = bas.distributed.render_by_time(x, 512, 32, traj, bank, gather=True)
if rank == 0:
    ref = bas.make_signal_move_2d(x, 512, 32, traj, bank)
    out['by_time_rel_l2_vs_one_gpu'] = rel(full, ref.astype(np.float64))
    out['by_time_peak'] = float(np.abs(full).max())
    out['by_time_shape'] = list(full.shape)
    ok = (out['by_source_peer_rel_l2_vs_one_gpu'] < 1e-6 and out['by_source_nccl_rel_l2_vs_one_gpu'] < 1e-6 and out['by_source_peer_same_on_every_rank'] and
          out['by_source_sourceless_rank_rel_l2'] < 1e-6 and out['stream_replicated_rel_l2_vs_one_gpu'] < 1e-6 and
          out['stream_sharded_rel_l2_vs_one_gpu'] < 1e-6 and out['stream_same_on_every_rank'] and out['stream_error_raises_on_every_rank'] and out['by_time_rel_l2_vs_one_gpu'] < 1e-6 and abs(out['by_time_peak'] - 1) < 1e-6)
    out['world'] = world
    out['ok'] = bool(ok)
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (rank != 0 or out['ok']) else 1)
