#!/usr/bin/env python
"""cProfile of the host side of a small training step (the launch-bound regime): where the microseconds go."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_n10"])
dev = torch.device("cuda:0")
layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
x = x_mem.permute(0, 2, 1)
dy = dy_mem.permute(0, 2, 1) if w["concat"] else dy_mem.permute(0, 2, 1).contiguous()


def step():
    for p in layer.parameters():
        p.grad = None
    xg = x.detach().requires_grad_(True)
    layer.addGSO(S)
    layer(xg).backward(dy)


for _ in range(20):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(200):
    step()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) / 200 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
