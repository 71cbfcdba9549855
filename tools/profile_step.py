#!/usr/bin/env python
"""One forward (or forward + backward) of the bench workload inside cudaProfilerStart/Stop, for ncu:

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof_fwd python tools/profile_step.py --what fwd
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="fwd", choices=["fwd", "train"])
ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
w = dict(bench.WORKLOADS[args.workload])
if args.batch:
    w["B"] = args.batch
dev = torch.device("cuda:0")
layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
x = x_mem.permute(0, 2, 1)
dy = dy_mem.permute(0, 2, 1)


def step():
    if args.what == "fwd":
        with torch.no_grad():
            layer.addGSO(S)
            return layer(x)
    for p in layer.parameters():
        p.grad = None
    xg = x.detach().requires_grad_(True)
    layer.addGSO(S)
    y = layer(xg)
    y.backward(dy)
    return y


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
