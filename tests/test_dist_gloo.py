"""World-size-2 gloo test of the multi-GPU host logic (batch sharding, gradient all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magat_pathplanning_b200.dist import shard_bounds


def test_shard_bounds_cover_batch_exactly():
    for batch in (1, 2, 7, 512, 1024, 1025):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(batch, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == batch
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magat_pathplanning_b200 import GraphFilterBatchAttentional
        from magat_pathplanning_b200.dist import allreduce_gradients, shard_batch
        torch.manual_seed(7)
        layer = GraphFilterBatchAttentional(16, 16, 3, 2, 1, True, concatenate=True, attentionMode="KeyQuery")
        # every rank fakes a local gradient = (rank + 1) * ones for the params KeyQuery trains
        for name, p in layer.named_parameters():
            if name in ("mixer", "weight_bias"):
                continue                                  # stay None, like the reference in KeyQuery mode
            p.grad = torch.full_like(p, float(rank + 1))
        n = allreduce_gradients(layer.parameters())
        ok = n == sum(p.numel() for k, p in layer.named_parameters() if k not in ("mixer", "weight_bias"))
        for name, p in layer.named_parameters():
            if name in ("mixer", "weight_bias"):
                ok = ok and p.grad is None
            else:
                ok = ok and bool(torch.all(p.grad == (1 + world) / 2.0))      # mean of 1..world
        # uneven shards: weighted combination = gradient of the global-batch mean; a parameter only one rank has a
        # gradient for (an empty shard elsewhere) still reduces like-for-like
        lin = torch.nn.Linear(3, 2)
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.ones_like(lin.bias) if rank == 0 else None
        allreduce_gradients(lin.parameters(), local_weight=float(3 if rank == 0 else 1), uniform=False)
        ok = ok and bool(torch.allclose(lin.weight.grad, torch.full_like(lin.weight, (3 * 1 + 1 * 2) / 4.0)))
        ok = ok and lin.bias.grad is not None and bool(torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 3 / 4.0)))
        full = torch.arange(10 * 3).reshape(10, 3)
        mine = shard_batch(full, rank, world)
        gathered = [torch.zeros(5, 3, dtype=full.dtype) for _ in range(world)]
        dist.all_gather(gathered, mine.contiguous())
        ok = ok and torch.equal(torch.cat(gathered), full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}
