#!/usr/bin/env python
"""Per-kernel times of one training step of a bench workload (device events between launches inside the library):
    python tools/time_train_kernels.py [--workload c4_n1000] [--filter tap_bwd]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from magat_pathplanning_b200 import _cabi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
ap.add_argument("--filter", default="")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
w = dict(bench.WORKLOADS[args.workload])
dev = torch.device("cuda:0")
layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
x, dy = x_mem.permute(0, 2, 1), dy_mem.permute(0, 2, 1)


def step():
    for p in layer.parameters():
        p.grad = None
    xg = x.detach().requires_grad_(True)
    layer.addGSO(S)
    layer(xg).backward(dy)


for _ in range(3):
    step()
torch.cuda.synchronize()
L = _cabi.lib()
L.magat_profile_enable(1)
for _ in range(args.steps):
    step()
torch.cuda.synchronize()
rec = _cabi.profile_collect()
L.magat_profile_enable(0)
tot = {}
for name, _n, ms in rec:
    tot[name] = tot.get(name, 0.0) + ms / args.steps
print(" ".join(f"{k}={v:.4f}" for k, v in tot.items() if args.filter in k), f"total={sum(tot.values()):.4f}")
