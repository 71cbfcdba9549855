"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/gat_golden.npz

Imports /root/reference/utils/graphUtils/graphML.py through oracle/ref_loader.py, runs
``GraphFilterBatchAttentional`` (graphML.py:4506) forward and autograd backward on CPU in
fp32 for a fixed list of seeded cases and stores inputs, parameters, outputs, the attention
tensor and all gradients.  The reference has no tests of its own, so these
files are what pins both ``oracle/gat_oracle.py`` and the CUDA path.
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

# name, mode, concat, G, F, K, P, B, N, extras
CASES = [
    dict(name="kq_concat_n10", mode="KeyQuery", concat=True, G=16, F=16, K=3, P=4, B=3, N=10),
    dict(name="kq_mean_c1", mode="KeyQuery", concat=False, G=128, F=128, K=2, P=1, B=1, N=10),
    dict(name="kq_concat_c2", mode="KeyQuery", concat=True, G=128, F=128, K=3, P=4, B=4, N=10),
    dict(name="gm_concat_fneg", mode="GAT_modified", concat=True, G=16, F=24, K=3, P=4, B=2, N=10,
         wb_std=0.1, x_signed=True),
    dict(name="gm_mean_n37", mode="GAT_modified", concat=False, G=32, F=8, K=2, P=2, B=2, N=37,
         wb_std=0.1),
    dict(name="kq_k1", mode="KeyQuery", concat=True, G=8, F=8, K=1, P=2, B=2, N=5),
    dict(name="kq_n1", mode="KeyQuery", concat=True, G=8, F=8, K=2, P=2, B=2, N=1),
    dict(name="gm_n2", mode="GAT_modified", concat=False, G=8, F=4, K=3, P=3, B=2, N=2, gso="full"),
    dict(name="kq_full_gso", mode="KeyQuery", concat=True, G=16, F=16, K=3, P=2, B=2, N=12, gso="full"),
    dict(name="gm_full_gso", mode="GAT_modified", concat=True, G=16, F=16, K=2, P=2, B=1, N=12, gso="full",
         wb_std=0.1),
    dict(name="kq_weird_gso", mode="KeyQuery", concat=True, G=16, F=16, K=3, P=2, B=2, N=9, gso="weird"),
    dict(name="gm_weird_gso", mode="GAT_modified", concat=False, G=16, F=16, K=3, P=2, B=2, N=9, gso="weird",
         wb_std=0.1, x_signed=True),
    dict(name="kq_pad", mode="KeyQuery", concat=True, G=16, F=16, K=3, P=2, B=2, N=11, Nin=8),
    dict(name="gm_pad_mean", mode="GAT_modified", concat=False, G=16, F=8, K=2, P=2, B=2, N=11, Nin=7,
         wb_std=0.1),
    dict(name="kq_f64_gso", mode="KeyQuery", concat=False, G=16, F=16, K=2, P=4, B=2, N=10, s_dtype="float64"),
    dict(name="kq_nobias", mode="KeyQuery", concat=True, G=16, F=16, K=2, P=2, B=2, N=10, bias=False),
    dict(name="kq_b32p4_n100", mode="KeyQuery", concat=False, G=32, F=32, K=2, P=4, B=2, N=100, width=50),
    dict(name="kq_n130_g128", mode="KeyQuery", concat=True, G=128, F=128, K=3, P=4, B=1, N=130, width=40),
    dict(name="gm_n70_g64", mode="GAT_modified", concat=True, G=64, F=64, K=3, P=2, B=2, N=70, width=25,
         wb_std=0.1),
    dict(name="kq_odd_dims", mode="KeyQuery", concat=True, G=20, F=20, K=4, P=3, B=2, N=33, width=14),
    dict(name="gm_odd_dims", mode="GAT_modified", concat=True, G=12, F=20, K=4, P=3, B=2, N=33, width=14,
         wb_std=0.1, x_signed=True),
]


def make_gso(case, gen):
    B, N = case["B"], case["N"]
    kind = case.get("gso", "geometric")
    if kind == "full":                       # decentralplanner_GAT.py:274-275 (ones incl. the diagonal)
        S = torch.ones(B, 1, N, N)
    elif kind == "weird":                    # asymmetric, negative weights, NaN, sub-tolerance values
        S = (torch.rand(B, 1, N, N, generator=gen) < 0.3).float()
        S = S * (torch.randn(B, 1, N, N, generator=gen))
        S[:, :, 0, :] = 0.0                  # an empty row
        S[:, :, :, 1] = 0.0                  # an empty column
        S[:, :, 2, 3] = float("nan")         # NaN compares false -> no edge
        S[:, :, 3, 4] = 5e-10                # below zeroTolerance -> no edge
        S[:, :, 4, 5] = -2e-9                # |.| above zeroTolerance -> edge
        S[:, :, 5, 5] = 1.0                  # a self loop
    else:
        S = random_geometric_gso(B, N, width=case.get("width"), generator=gen)
    return S.to(getattr(torch, case.get("s_dtype", "float32")))


def run_case(gml, case):
    gen = torch.Generator().manual_seed(1337 + sum(map(ord, case["name"])))
    torch.manual_seed(20261017 + sum(map(ord, case["name"])))
    G, F, K, P, B, N = (case[k] for k in "GFKPBN")
    Nin = case.get("Nin", N)
    layer = gml.GraphFilterBatchAttentional(G, F, K, P, 1, case.get("bias", True),
                                            concatenate=case["concat"], attentionMode=case["mode"])
    if case.get("wb_std"):
        with torch.no_grad():
            layer.weight_bias.normal_(0.0, case["wb_std"])
    S = make_gso(case, gen)
    xm = torch.randn(B, Nin, G, generator=gen)
    if not case.get("x_signed"):
        xm = torch.relu(xm)
    x = xm.permute(0, 2, 1).clone().requires_grad_(True)      # [B,G,Nin]
    layer.addGSO(S)
    y = layer(x)
    dy = torch.randn(y.shape, generator=gen)
    y.backward(dy)
    out = {"x": x.detach(), "S": S, "dy": dy, "y": y.detach(), "aij": torch.from_numpy(layer.aij)}
    none_grads = []
    for pname, p in layer.named_parameters():
        out["param." + pname] = p.detach()
        if p.grad is None:
            none_grads.append(pname)
        else:
            out["grad." + pname] = p.grad
    out["grad.x"] = x.grad
    meta = dict(case)
    meta["none_grads"] = none_grads
    meta["returnAttentionGSO_shape"] = list(layer.returnAttentionGSO().shape)
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
    path = os.path.join(HERE, "gat_golden.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(CASES)} cases, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
