#!/usr/bin/env python
"""Which kernels a configuration runs at scale, and for how long: looks for shapes that fall off the fast paths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from magat_pathplanning_b200 import _cabi  # noqa: E402

base = dict(B=128, N=1000, width=200, G=128, K=3, P=4, concat=True, mode="KeyQuery")
variants = {
    "base": {}, "gm_mean": dict(mode="GAT_modified", concat=False), "p2": dict(P=2), "p1": dict(P=1), "k1": dict(K=1),
    "k2": dict(K=2), "p1_mean_k2": dict(P=1, K=2, concat=False), "g64": dict(G=64), "g32_mean_k2": dict(G=32, K=2, concat=False),
    "g256_p1": dict(G=256, P=1, K=2), "p3": dict(P=3), "k4": dict(K=4),
}
dev = torch.device("cuda:0")
L = _cabi.lib()
for name, kw in variants.items():
    w = dict(base)
    w.update(kw)
    try:
        layer, S, x_mem, dy_mem = bench.make_problem(w, dev, bench.SEED)
        x = x_mem.permute(0, 2, 1)
        dy = dy_mem.permute(0, 2, 1) if w["concat"] else dy_mem.permute(0, 2, 1).contiguous()

        def step():
            for p in layer.parameters():
                p.grad = None
            xg = x.detach().requires_grad_(True)
            layer.addGSO(S)
            layer(xg).backward(dy)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        L.magat_profile_enable(1)
        step()
        torch.cuda.synchronize()
        rec = _cabi.profile_collect()
        L.magat_profile_enable(0)
        tot = sum(ms for _n, _c, ms in rec)
        slow = sorted(rec, key=lambda r: -r[2])[:4]
        print(f"{name:14s} total {tot:7.3f} ms  " + "  ".join(f"{n}={ms:.3f}" for n, _c, ms in slow))
        del layer, S, x_mem, dy_mem
    except Exception as exc:  # noqa: BLE001
        print(f"{name:14s} ERROR {str(exc)[-200:]}")
