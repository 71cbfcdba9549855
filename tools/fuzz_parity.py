#!/usr/bin/env python
"""Random configurations (mode, concat, G, F, K, P, B, N, bias, GSO dtype) through path="auto", forward and every
gradient against the CPU oracle.  Looks for dispatch mismatches (a *_supported predicate promising what a kernel cannot
take), not for performance.      python tools/fuzz_parity.py [n_cases] [seed]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import gat_oracle as orc  # noqa: E402
from magat_pathplanning_b200 import GraphFilterBatchAttentional  # noqa: E402
from magat_pathplanning_b200 import graphML as ours  # noqa: E402

PARAMS = ("mixer", "weight_bias", "filterWeight", "bias", "weight")


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(n_cases, seed, verbose=True):
    """Returns the descriptions of the cases that failed."""
    rng = random.Random(seed)
    dev = torch.device("cuda:0")
    failures = []
    for it in range(n_cases):
      mode = rng.choice(["KeyQuery", "GAT_modified"])
      concat = rng.random() < 0.6
      G = rng.choice([16, 32, 64, 128, 128, 128, 256])
      F = G if mode == "KeyQuery" else rng.choice([G, 128, 64, 32])
      K, P = rng.randint(1, 5), rng.randint(1, 5)
      B, N = rng.randint(1, 6), rng.choice([1, 3, 10, 31, 32, 33, 64, 100, 130, 200, 257, 500])
      bias = rng.random() < 0.8
      s_dtype = rng.choice([torch.float32, torch.float32, torch.float64])
      tag = f"{mode} concat={concat} G={G} F={F} K={K} P={P} B={B} N={N} bias={bias} {s_dtype}"
      try:
          gen = torch.Generator().manual_seed(1000 + it)
          params = orc.init_params(G, F, K, P, mode=mode, bias=bias, generator=gen, weight_bias_std=0.1)
          S = orc.random_geometric_gso(B, N, generator=gen).to(s_dtype)
          x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
          _, _, pre = orc.gat_layer_forward(x, S, params, mode=mode, concatenate=concat, return_pre=True)
          pre_out = pre.reshape(B, P * F, N) if concat else pre.mean(dim=1)
          dy = torch.randn(pre_out.shape, generator=gen) * (pre_out.abs() > 1e-3)
          y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(x, S, params, dy, mode=mode, concatenate=concat)
          layer = GraphFilterBatchAttentional(G, F, K, P, 1, bias, concatenate=concat, attentionMode=mode)
          with torch.no_grad():
              for k in PARAMS:
                  if params.get(k) is not None and getattr(layer, k, None) is not None:
                      getattr(layer, k).copy_(params[k])
          layer = layer.to(dev)
          layer.addGSO(S.to(dev))
          xd = x.to(dev).requires_grad_(True)
          y = layer(xd)
          errs = {"y": rel_err(y, y_ref), "aij": float((torch.from_numpy(layer.aij) - aij_ref).abs().max())}
          y.backward(dy.to(dev))
          errs["dx"] = rel_err(xd.grad, g_ref["x"])
          gmax = max(float(v.abs().max()) for v in g_ref.values() if v is not None)
          for k in PARAMS:
              if g_ref.get(k) is not None:
                  if float(g_ref[k].abs().max()) < 1e-6 * gmax:      # mathematically zero (GAT_modified's weight_bias): noise
                      errs["d" + k] = float(getattr(layer, k).grad.abs().max()) / gmax
                  else:
                      errs["d" + k] = rel_err(getattr(layer, k).grad, g_ref[k])
          worst = max(errs.values())
          if worst > 1e-4 or list(y.stride()) != list(y_ref.stride()):
              failures.append("MISMATCH " + tag + " " + str({k: f"{v:.1e}" for k, v in errs.items()}))
              if verbose:
                  print(failures[-1])
      except Exception as exc:  # noqa: BLE001
          failures.append("ERROR " + tag + " " + str(exc)[-160:])
          if verbose:
              print(failures[-1])
    return failures


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print(f"{n} cases, {len(bad)} bad")
