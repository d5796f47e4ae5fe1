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


def evaluate_sharded(fn, x: torch.Tensor, rank: int, world_size: int, group=None, out: torch.Tensor = None):
    """fn(x) for a row-wise function fn, evaluated as world_size contiguous slices -- rank r computes rows
    [r * chunk, (r + 1) * chunk) with chunk = ceil(n / world_size) -- and all-gathered, so every rank gets the full result
    having done 1/world_size of the work (the occupancy-grid refresh of data-parallel training: every rank draws the same cells
    and jitter, the per-row result does not depend on which rank computed it).  `out` (optional): a buffer of at least
    chunk * world_size elements to gather into.  Returns a view of n elements."""
    n = x.shape[0]
    chunk = (n + world_size - 1) // world_size
    lo = min(rank * chunk, n)
    hi = min(lo + chunk, n)
    full = out[:chunk * world_size] if out is not None else torch.empty(chunk * world_size, dtype=torch.float32, device=x.device)
    mine = torch.zeros(chunk, dtype=torch.float32, device=x.device)
    if hi > lo:
        mine[:hi - lo] = fn(x[lo:hi]).reshape(-1)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


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


class PeerGradients:
    """Gradient buffers of all ranks mapped into every process (torch symmetric memory over NVLink / NVSwitch) for the fused
    all-reduce + Adam kernel (angio_adam_step_allreduce).  Layout of each rank's allocation, in 4-byte words:
    [gradient, even steps (n_alloc) | gradient, odd steps (n_alloc) | flags[world] (uint32 step tags written by the peers)].
    torch.distributed only does the plumbing (allocation + handle exchange); the data path is our kernel."""

    def __init__(self, n_floats, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("PeerGradients supports up to 16 ranks (one NVSwitch domain)")
        self.n = int(n_floats)
        self.n_alloc = (self.n + 63) // 64 * 64
        words = 2 * self.n_alloc + 64
        self.buf = symm.empty(words, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self.grad = [self.buf[k * self.n_alloc:k * self.n_alloc + self.n] for k in range(2)]
        self.flags = self.buf[2 * self.n_alloc:2 * self.n_alloc + 64].view(torch.int32)
        import ctypes
        arr = ctypes.c_void_p * self.world
        self.peer_grad_ptrs = [arr(*[b + 4 * k * self.n_alloc for b in base]) for k in range(2)]
        self.peer_flag_ptrs = arr(*[b + 4 * 2 * self.n_alloc for b in base])
        self.tag = 0
        # [ns spent waiting for peer tags, calls, longest wait, -] accumulated by the fused all-reduce + Adam kernel
        self.wait_stats = torch.zeros(4, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # every rank's buffer is zeroed before anyone signals

    def next_buffer(self):
        """(gradient buffer of the coming step, its tag).  Tags start at 1; the buffer alternates with the tag's parity."""
        self.tag += 1
        return self.grad[self.tag & 1], self.tag
