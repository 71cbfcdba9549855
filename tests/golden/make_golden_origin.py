"""Golden vectors of the GAT_origin ablation from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_origin.py      # writes tests/golden/origin_golden.npz

``GraphFilterBatchAttentional_Origin`` (graphML.py:4175-4339 over :1939-2005 and :964-1070), forward and autograd
backward on CPU in fp32, seeded.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.gat_oracle import random_geometric_gso  # noqa: E402  (input generator only)
from oracle.ref_loader import load_reference_graphml  # noqa: E402

CASES = [
    dict(name="or_concat_n10", G=16, F=16, K=3, P=2, B=3, N=10, concat=True),
    dict(name="or_mean_n10", G=16, F=16, K=2, P=4, B=2, N=10, concat=False),
    dict(name="or_g128_n12", G=128, F=128, K=3, P=4, B=2, N=12, concat=True),
    dict(name="or_fneg", G=24, F=12, K=3, P=2, B=2, N=9, concat=True),        # F != G: the reshape of W scrambles
    dict(name="or_k1", G=8, F=8, K=1, P=2, B=2, N=6, concat=True),
    dict(name="or_selfloop_cancel", G=16, F=16, K=2, P=1, B=2, N=8, concat=True, gso="minus_diag"),
    dict(name="or_pad", G=16, F=16, K=2, P=2, B=2, N=11, Nin=8, concat=False),
    dict(name="or_nobias_f64", G=16, F=16, K=2, P=2, B=2, N=10, concat=True, bias=False, s_dtype="float64"),
    dict(name="or_g128_n130", G=128, F=128, K=3, P=4, B=1, N=130, width=40, concat=True),
]


def make_gso(case, gen):
    B, N = case["B"], case["N"]
    S = random_geometric_gso(B, N, width=case.get("width"), generator=gen)
    if case.get("gso") == "minus_diag":          # S + I is zero on the diagonal of node 2: no self loop there (:1019-1024)
        S[:, :, 2, 2] = -1.0
    return S.to(getattr(torch, case.get("s_dtype", "float32")))


def run_case(gml, case):
    gen = torch.Generator().manual_seed(777 + sum(map(ord, case["name"])))
    torch.manual_seed(20261019 + sum(map(ord, case["name"])))
    G, F, K, P, B, N = (case[k] for k in ("G", "F", "K", "P", "B", "N"))
    Nin = case.get("Nin", N)
    layer = gml.GraphFilterBatchAttentional_Origin(G, F, K, P, 1, case.get("bias", True), concatenate=case["concat"])
    with torch.no_grad():                        # taps of order one so that every tap matters in the comparison
        layer.filterWeight.copy_(torch.randn(1, K, generator=gen))
    S = make_gso(case, gen)
    x = torch.relu(torch.randn(B, Nin, G, generator=gen)).permute(0, 2, 1).clone().requires_grad_(True)
    layer.addGSO(S)
    y = layer(x)
    dy = torch.randn(y.shape, generator=gen)
    y.backward(dy)
    out = {"x": x.detach(), "S": S, "dy": dy, "y": y.detach(), "grad.x": x.grad,
           "aij": torch.from_numpy(layer.aij)}
    for pname, p in layer.named_parameters():
        out["param." + pname] = p.detach()
        if p.grad is not None:
            out["grad." + pname] = p.grad
    meta = dict(case)
    meta["y_stride"] = list(y.stride())
    meta["grads"] = [n for n, p in layer.named_parameters() if p.grad is not None]
    return out, meta


def main():
    gml = load_reference_graphml()
    torch.set_num_threads(1)
    blob, metas = {}, []
    for case in CASES:
        out, meta = run_case(gml, case)
        for k, v in out.items():
            blob[f"{case['name']}/{k}"] = v.contiguous().numpy()
        metas.append(meta)
    blob["__meta__"] = np.frombuffer(json.dumps(metas).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "origin_golden.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(CASES)} cases, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
