"""Host <-> device copy bandwidth with every rank copying at once (run under torchrun): separates the PCIe links
from the host side (memory bandwidth / IOMMU of the VM) as the limit of the e2e numbers at N > 1.
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py"""
import os, sys, json, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
mb = 256
host = torch.empty(mb << 18, dtype=torch.float32, pin_memory=True)
host.fill_(1.0)
devt = torch.empty_like(host, device='cuda')
res = {}
for name, fn in (('h2d', lambda: devt.copy_(host, non_blocking=True)), ('d2h', lambda: host.copy_(devt, non_blocking=True))):
    for solo in (True, False):
        if solo and world > 1:
            # one rank at a time
            gbs = 0.0
            for r in range(world):
                dist.barrier()
                if r == rank:
                    fn(); torch.cuda.synchronize()
                    t0 = time.perf_counter(); fn(); fn(); torch.cuda.synchronize(); gbs = 2 * host.numel() * 4 / (time.perf_counter() - t0) / 1e9
                dist.barrier()
            res[name + '_alone_gbs'] = gbs
        else:
            if world > 1:
                dist.barrier()
            fn(); torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter(); fn(); fn(); torch.cuda.synchronize(); gbs = 2 * host.numel() * 4 / (time.perf_counter() - t0) / 1e9
            res[name + ('_all_ranks_at_once_gbs' if world > 1 else '_gbs')] = gbs
every = [res]
if world > 1:
    every = [None] * world
    dist.all_gather_object(every, res)
if rank == 0:
    keys = sorted(every[0])
    print(json.dumps({'world': world, 'per_rank': {k: [round(e[k], 1) for e in every] for k in keys},
                      'aggregate': {k: round(sum(e[k] for e in every), 1) for k in keys if 'at_once' in k or world == 1}}))
if world > 1:
    dist.destroy_process_group()
