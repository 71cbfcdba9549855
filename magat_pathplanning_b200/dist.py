"""Multi-GPU plumbing of the path: planning instances are independent graphs, so the batch
is sharded contiguously across ranks with NO data-path collective; training adds one all-reduce of the
parameter gradients (NCCL over NVLink on GPUs; any torch.distributed backend works, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `batch` instances: ranks < batch % world get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        average: bool = True) -> int:
    """Sum (then average) the gradients of `params` across ranks through ONE flat buffer (the layer holds
    ~1 MB of parameters: latency bound, so a single collective).  Parameters whose grad is None keep None
    -- KeyQuery leaves mixer / weight_bias untouched and Adam must not decay them.
    Returns the number of elements reduced."""
    ps = [p for p in params if p.grad is not None]
    if not ps or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum(p.grad.numel() for p in ps)
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off
