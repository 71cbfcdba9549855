#!/usr/bin/env python
"""One fused forward (path='fused') of the bench workload inside cudaProfilerStart/Stop, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD])
dev = torch.device("cuda:0")
layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
layer.path = "fused"
layer.max_degree = 16
layer.fused_team = int(os.environ.get("MAGAT_TEAM", "8"))
x = x_mem.permute(0, 2, 1)
def step():
    with torch.no_grad():
        layer.addGSO(S)
        return layer(x)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
