#!/usr/bin/env python
"""Back-to-back training steps of the bench workload (no sync in between); prints the first failure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
w = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
if len(sys.argv) > 2:
    w["B"] = int(sys.argv[2])
dev = torch.device("cuda:0")
layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
x, dy = x_mem.permute(0, 2, 1), dy_mem.permute(0, 2, 1)
try:
    for i in range(n):
        for p in layer.parameters():
            p.grad = None
        xg = x.detach().requires_grad_(True)
        layer.addGSO(S)
        y = layer(xg)
        y.backward(dy)
    torch.cuda.synchronize()
    print("ok", n, "steps", float(layer.filterWeight.grad.abs().sum()), float(xg.grad.abs().sum()))
except Exception as e:
    print("FAILED at step", i, ":", str(e)[-300:])
