#!/usr/bin/env python
"""Per-kernel times of the inference step with the fused action head (forward_actions) and of the two-step route."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from magat_pathplanning_b200 import _cabi  # noqa: E402

w = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
dev = torch.device("cuda:0")
layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
x = x_mem.permute(0, 2, 1)
head = torch.nn.Linear(w["P"] * 128, 5).to(dev)
L = _cabi.lib()
for name in ("fused", "two-step"):
    def step():
        with torch.no_grad():
            layer.addGSO(S)
            if name == "fused":
                return layer.forward_actions(x, head)
            return head(layer(x).permute(0, 2, 1).reshape(-1, w["P"] * 128))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    L.magat_profile_enable(1)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    rec = _cabi.profile_collect()
    L.magat_profile_enable(0)
    print(name, " ".join(f"{n}={ms / 5:.4f}" for n, _c, ms in rec), f"total={sum(ms for _n, _c, ms in rec) / 5:.4f}")
