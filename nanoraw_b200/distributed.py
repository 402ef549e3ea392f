"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous and for combining
per-rank counters.  The data path has NO collective: reads are independent
(resquiggle.py:318-401 keeps no cross-read state), so each rank re-squiggles its own contiguous
range of reads (balanced by raw samples) and only scalars are reduced."""
import os

import numpy as np


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def init(backend=None):
    """Initialise torch.distributed when launched with more than one process."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kwargs['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def shard_batch(batch, rank, world):
    """this rank's reads of a host ReadBatch"""
    return batch.shard(rank, world)


def combine(max_values, sum_values, device='cpu'):
    """max over ranks of `max_values` (timings) and sum over ranks of `sum_values` (work
    counters): the only communication of a multi-GPU run."""
    import torch
    import torch.distributed as dist
    mx = torch.tensor(list(max_values), dtype=torch.float64, device=device)
    sm = torch.tensor(list(sum_values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx.tolist(), sm.tolist()


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def gather_statuses(status, device='cpu'):
    """all ranks' per-read status arrays concatenated in rank order (rank 0 writes the failed
    read summary); statuses are padded to the longest shard."""
    import torch
    import torch.distributed as dist
    status = np.asarray(status, dtype=np.int32)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return status
    world = dist.get_world_size()
    n = torch.tensor([len(status)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    longest = int(max(int(s.item()) for s in sizes))
    mine = torch.full((longest,), -1, dtype=torch.int32, device=device)
    mine[:len(status)] = torch.from_numpy(status).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return np.concatenate([p.cpu().numpy()[:int(s.item())] for p, s in zip(parts, sizes)])
