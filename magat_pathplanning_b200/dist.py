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
                        average: bool = True, local_weight: Optional[float] = None, uniform: bool = True) -> int:
    """Combine the gradients of `params` across ranks through ONE flat buffer (the layer holds ~1 MB of parameters:
    latency bound, so a single collective).  Nothing here synchronises the host unless ``uniform=False``.

    * ``average=True, local_weight=None``: plain mean over ranks -- right when every rank's loss is a mean over an
      equally sized shard.
    * ``local_weight=w`` (e.g. the number of instances of this rank's shard): sum_r w_r g_r / sum_r w_r, the gradient of
      the global-batch mean when `shard_bounds` hands out uneven shards (the weight rides in the same buffer and the
      division happens on the device).
    * ``uniform=True`` (default) is the caller's statement that every rank holds gradients for the SAME parameters
      (ordinary data parallelism).  With ``uniform=False`` the buffer is laid out over ALL parameters in iteration
      order with a per-parameter "have a gradient" flag, so ranks that disagree (an empty shard, a frozen parameter)
      still reduce like-for-like; the flags are read back (one host synchronisation).
    * A parameter whose grad is None (on every rank) keeps None: KeyQuery leaves mixer / weight_bias untouched and Adam
      must not decay them.

    Returns the number of gradient elements reduced."""
    ps = list(params)
    have = [p.grad is not None for p in ps]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum(p.grad.numel() for p, h in zip(ps, have) if h)
    world = dist.get_world_size(group)
    if uniform:
        ps = [p for p, h in zip(ps, have) if h]
        have = [True] * len(ps)
    if not ps:
        return 0
    dev, dt = ps[0].device, torch.float32
    w = 1.0 if local_weight is None else float(local_weight)
    pieces = [(p.grad.reshape(-1).to(dt) * w) if h else torch.zeros(p.numel(), dtype=dt, device=dev)
              for p, h in zip(ps, have)]
    n_flag = 0 if uniform else len(ps)
    tail = torch.full((n_flag + 1,), w, dtype=dt, device=dev)
    if n_flag:
        tail[:n_flag] = torch.tensor([1.0 if h else 0.0 for h in have], dtype=dt)
    flat = torch.cat(pieces + [tail])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if local_weight is not None:
        flat[:-1 - n_flag].div_(flat[-1].clamp_min(1e-30))
    elif average:
        flat[:-1 - n_flag].div_(world)
    flags = [1.0] * len(ps) if uniform else flat[-(n_flag + 1):-1].tolist()
    off = done = 0
    for p, h, f in zip(ps, have, flags):
        n = p.numel()
        if f > 0:
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if h:
                p.grad.copy_(g)
            else:
                p.grad = g.clone()
            done += n
        off += n
    return done
