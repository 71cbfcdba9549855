"""SURVEY 8f row f4, first ablation mode: GAT_origin (graphML.py:4175-4339, :1939-2005, :964-1070) -- oracle against the
golden vectors of the unmodified reference (CPU), CUDA path against the golden vectors (GPU).  Tolerance 1e-4 max-norm
relative."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gat_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "golden", "origin_golden.npz")
TOL = 1e-4
PARAMS = ("mixer", "weight", "filterWeight", "bias")


def cases():
    z = np.load(PATH)
    return [m["name"] for m in json.loads(bytes(z["__meta__"]).decode())]


def load(name):
    z = np.load(PATH)
    meta = {m["name"]: m for m in json.loads(bytes(z["__meta__"]).decode())}[name]
    d = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/")}
    return d, meta


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", cases())
def test_oracle_matches_reference_golden(name):
    d, meta = load(name)
    params = {k: d["param." + k].clone().requires_grad_(True) for k in PARAMS if "param." + k in d}
    x = d["x"].clone().requires_grad_(True)
    y, aij = orc.origin_layer_forward(x, d["S"], params, concatenate=meta["concat"])
    assert y.shape == d["y"].shape
    assert rel_err(y, d["y"]) < 2e-6
    assert float((aij.detach() - d["aij"]).abs().max()) < 2e-6
    y.backward(d["dy"])
    assert rel_err(x.grad, d["grad.x"]) < 2e-5
    for k, p_ in params.items():
        if k in meta["grads"]:
            assert rel_err(p_.grad, d["grad." + k]) < 2e-5, k
        else:
            assert p_.grad is None or float(p_.grad.abs().max()) == 0.0, k


def test_module_surface_matches_reference():
    from magat_pathplanning_b200.graphML import GraphFilterBatchAttentional_Origin as Ours
    m = Ours(16, 24, 3, 2, 1, True)
    assert [k for k, _ in m.named_parameters()] == ["mixer", "weight", "filterWeight", "bias"]
    assert tuple(m.mixer.shape) == (2, 1, 48) and tuple(m.weight.shape) == (2, 1, 24, 16)
    assert tuple(m.filterWeight.shape) == (1, 3) and tuple(m.bias.shape) == (24, 1)
    assert "no GSO stored" in repr(m)
    from oracle.ref_loader import load_reference_graphml, reference_available
    if reference_available():
        import inspect
        import magat_pathplanning_b200.graphML as ours
        gml = load_reference_graphml()
        ref = gml.GraphFilterBatchAttentional_Origin(16, 24, 3, 2, 1, True)
        assert list(ref.state_dict()) == list(m.state_dict())
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert repr(ref) == repr(m)
        for fn in ("graphAttentionLSIGFBatch_Origin", "learnAttentionGSOBatch_origin"):
            assert str(inspect.signature(getattr(ours, fn))) == str(inspect.signature(getattr(gml, fn))), fn


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["simt", "auto"])
@pytest.mark.parametrize("name", cases())
def test_cuda_matches_reference_golden(name, path):
    from magat_pathplanning_b200.graphML import GraphFilterBatchAttentional_Origin as Ours
    d, meta = load(name)
    dev = torch.device("cuda:0")
    layer = Ours(meta["G"], meta["F"], meta["K"], meta["P"], 1, meta.get("bias", True), concatenate=meta["concat"])
    with torch.no_grad():
        for k in PARAMS:
            if "param." + k in d:
                getattr(layer, k).copy_(d["param." + k])
    layer = layer.to(dev)
    layer.path = path
    x = d["x"].to(dev).requires_grad_(True)
    layer.addGSO(d["S"].to(dev))
    y = layer(x)
    assert y.shape == d["y"].shape and list(y.stride()) == meta["y_stride"]
    assert rel_err(y, d["y"]) < TOL
    assert float((torch.from_numpy(layer.aij) - d["aij"]).abs().max()) < TOL
    assert rel_err(torch.from_numpy(layer.returnAttentionGSO()), d["aij"].mean(dim=1)) < TOL
    y.backward(d["dy"].to(dev))
    assert rel_err(x.grad, d["grad.x"]) < TOL
    for k in PARAMS:
        p_ = getattr(layer, k, None)
        if p_ is None:
            continue
        if k in meta["grads"]:
            assert rel_err(p_.grad, d["grad." + k]) < TOL, k
        else:
            assert p_.grad is None or float(p_.grad.abs().max()) == 0.0, k


@pytest.mark.gpu
def test_functionals_and_rebinding():
    import magat_pathplanning_b200.graphML as ours
    d, meta = load("or_concat_n10")
    dev = torch.device("cuda:0")
    h, a, W, b = (d["param." + k].to(dev) for k in ("filterWeight", "mixer", "weight", "bias"))
    with torch.no_grad():
        y, aij = ours.graphAttentionLSIGFBatch_Origin(h, d["x"].to(dev), a, W, d["S"].to(dev), b=b)
        aij2 = ours.learnAttentionGSOBatch_origin(d["x"].to(dev), a, W, d["S"].to(dev))
    B, N, P, F = meta["B"], meta["N"], meta["P"], meta["F"]
    assert y.shape == (B, P, F, N)
    y_cat = torch.relu(y).permute(0, 3, 1, 2).reshape(B, N, P * F).permute(0, 2, 1)
    assert rel_err(y_cat, d["y"]) < TOL
    assert float((aij.cpu() - d["aij"]).abs().max()) < TOL and float((aij2.cpu() - d["aij"]).abs().max()) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("s_dtype", [torch.float32, torch.float64])
def test_cuda_matches_oracle_larger_with_self_loop_bits(s_dtype):
    """Beyond the simulator's shape the self loops are set as bits on the scanned mask (magat_gso_self_loops) instead of
    materialising S + I: forward, attention and every gradient against the oracle, with diagonals that cancel the loop
    (-1), keep it (anything else, NaN excepted) and an isolated agent whose only edge is its loop."""
    from magat_pathplanning_b200.graphML import GraphFilterBatchAttentional_Origin as Ours
    dev = torch.device("cuda:0")
    G = F = 128
    K, P, B, N = 3, 4, 6, 100
    gen = torch.Generator().manual_seed(99)
    S = orc.random_geometric_gso(B, N, generator=gen).to(s_dtype)
    S[:, 0, 3, 3] = -1.0                 # S + I = 0: no self loop for agent 3
    S[:, 0, 4, 4] = 0.5
    S[0, 0, 5, 5] = float("nan")         # NaN + 1 is NaN: no edge
    S[:, 0, 7, :] = 0.0                  # agent 7 sends to nobody but itself
    layer = Ours(G, F, K, P, 1, True, concatenate=True)
    with torch.no_grad():
        layer.filterWeight.copy_(torch.randn(1, K, generator=gen))
    params = {k: getattr(layer, k).detach().clone().requires_grad_(True) for k in PARAMS}
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    xr = x.clone().requires_grad_(True)
    y_ref, aij_ref = orc.origin_layer_forward(xr, S, params, concatenate=True)
    dy = torch.randn(y_ref.shape, generator=gen) * (y_ref.detach() > 1e-3)
    y_ref.backward(dy)
    layer = layer.to(dev)
    layer.addGSO(S.to(dev))
    xd = x.to(dev).requires_grad_(True)
    y = layer(xd)
    assert rel_err(y, y_ref) < TOL
    assert float((torch.from_numpy(layer.aij) - aij_ref.detach()).abs().max()) < TOL
    y.backward(dy.to(dev))
    assert rel_err(xd.grad, xr.grad) < TOL
    for k in PARAMS:
        assert rel_err(getattr(layer, k).grad, params[k].grad) < TOL, k
