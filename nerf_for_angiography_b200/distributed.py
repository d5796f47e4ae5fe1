"""One-process-per-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in CPU tests).

Rays, views and volume slabs are independent, so the data path has exactly one exchange: the all-reduce of the flat
fp32 gradient buffer once per training step.  Everything else here is index arithmetic.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous, balanced [lo, hi) slice of range(n) owned by `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_gradient(flat_grad: torch.Tensor, local_batch: int, group=None):
    """Gradient of the GLOBAL-batch mean loss from per-rank gradients of LOCAL-batch means: sum over ranks of
    grad_r * (local_batch_r / global_batch).  With equal batches this is the plain average.  Returns the global batch."""
    rank, ws = world()
    if ws == 1:
        return local_batch
    n = torch.tensor([float(local_batch)], device=flat_grad.device)
    dist.all_reduce(n, group=group)
    flat_grad.mul_(float(local_batch) / float(n.item()))
    dist.all_reduce(flat_grad, group=group)
    return int(n.item())


def gather_concat(local: torch.Tensor, group=None, dst: int = 0):
    """Concatenate per-rank tensors (ragged along dim 0) on rank `dst`; other ranks get None."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = torch.zeros((max(sizes),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
