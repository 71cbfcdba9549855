"""Golden vectors of the non-attentional graph filter from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_lsigf.py      # writes tests/golden/lsigf_golden.npz

``GraphFilterBatch`` / ``BatchLSIGF`` (graphML.py:5485-5700), forward and autograd backward on CPU in fp32, seeded.
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
    dict(name="gf_n10", G=16, F=24, K=3, B=3, N=10),
    dict(name="gf_g128_n10", G=128, F=128, K=3, B=4, N=10),
    dict(name="gf_k1", G=8, F=8, K=1, B=2, N=5),
    dict(name="gf_k4_odd", G=20, F=12, K=4, B=2, N=33, width=14, x_signed=True),
    dict(name="gf_full", G=16, F=16, K=3, B=2, N=12, gso="full"),
    dict(name="gf_weird", G=16, F=16, K=3, B=2, N=9, gso="weird", x_signed=True),
    dict(name="gf_pad", G=16, F=16, K=2, B=2, N=11, Nin=8),
    dict(name="gf_f64_gso", G=16, F=16, K=2, B=2, N=10, s_dtype="float64"),
    dict(name="gf_nobias", G=16, F=16, K=2, B=2, N=10, bias=False),
    dict(name="gf_g128_n130", G=128, F=128, K=3, B=1, N=130, width=40),
]


def make_gso(case, gen):
    B, N = case["B"], case["N"]
    kind = case.get("gso", "geometric")
    if kind == "full":
        S = torch.ones(B, 1, N, N) / N
    elif kind == "weird":                    # asymmetric, signed weights, an empty row / column, a tiny but non-zero entry
        S = (torch.rand(B, 1, N, N, generator=gen) < 0.3).float() * torch.randn(B, 1, N, N, generator=gen)
        S[:, :, 0, :] = 0.0
        S[:, :, :, 1] = 0.0
        S[:, :, 3, 4] = 5e-10                # BatchLSIGF multiplies by S itself: this IS an edge (unlike the attention mask)
        S[:, :, 5, 5] = 1.0
    else:
        S = random_geometric_gso(B, N, width=case.get("width"), generator=gen)
    return S.to(getattr(torch, case.get("s_dtype", "float32")))


def run_case(gml, case):
    gen = torch.Generator().manual_seed(4242 + sum(map(ord, case["name"])))
    torch.manual_seed(20261018 + sum(map(ord, case["name"])))
    G, F, K, B, N = (case[k] for k in "GFKBN")
    Nin = case.get("Nin", N)
    layer = gml.GraphFilterBatch(G, F, K, 1, case.get("bias", True))
    S = make_gso(case, gen)
    xm = torch.randn(B, Nin, G, generator=gen)
    if not case.get("x_signed"):
        xm = torch.relu(xm)
    x = xm.permute(0, 2, 1).clone().requires_grad_(True)
    layer.addGSO(S)
    y = layer(x)
    dy = torch.randn(y.shape, generator=gen)
    y.backward(dy)
    out = {"x": x.detach(), "S": S, "dy": dy, "y": y.detach(), "grad.x": x.grad}
    for pname, p in layer.named_parameters():
        out["param." + pname] = p.detach()
        out["grad." + pname] = p.grad
    meta = dict(case)
    meta["y_stride"] = list(y.stride())
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
    path = os.path.join(HERE, "lsigf_golden.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(CASES)} cases, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
