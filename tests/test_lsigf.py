"""GraphFilterBatch / BatchLSIGF (graphML.py:5485-5700) -- oracle against the golden vectors of the
unmodified reference (CPU), CUDA path against both (GPU).  Tolerance 1e-4 max-norm relative."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gat_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "golden", "lsigf_golden.npz")
TOL = 1e-4


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
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", cases())
def test_oracle_matches_reference_golden(name):
    d, meta = load(name)
    x, N = d["x"], meta["N"]
    if x.shape[2] < N:                                        # graphML.py:5674-5678 zero padding, :5684 slice
        x = torch.cat((x, torch.zeros(x.shape[0], x.shape[1], N - x.shape[2])), dim=2)
    dy = d["dy"]
    if dy.shape[2] < N:
        dy = torch.cat((dy, torch.zeros(dy.shape[0], dy.shape[1], N - dy.shape[2])), dim=2)
    y, g = orc.lsigf_fwd_bwd(x, d["S"], d["param.weight"], d.get("param.bias"), dy)
    Nin = d["x"].shape[2]
    assert rel_err(y[:, :, :Nin], d["y"]) < 2e-6
    assert rel_err(g["x"][:, :, :Nin], d["grad.x"]) < 2e-6
    assert rel_err(g["weight"], d["grad.weight"]) < 2e-6
    if "grad.bias" in d:
        assert rel_err(g["bias"], d["grad.bias"]) < 2e-6


def test_module_surface_matches_reference():
    from magat_pathplanning_b200 import GraphFilterBatch
    m = GraphFilterBatch(16, 24, 3, 1, True)
    assert [k for k, _ in m.named_parameters()] == ["weight", "bias"]
    assert tuple(m.weight.shape) == (24, 1, 3, 16) and tuple(m.bias.shape) == (24, 1)
    assert float(m.weight.abs().max()) <= 1.0 / (16 * 3) ** 0.5
    assert "no GSO stored" in repr(m)
    m.addGSO(torch.zeros(2, 1, 5, 5))
    assert m.N == 5 and "GSO stored" in repr(m)
    assert "bias" not in GraphFilterBatch(8, 8, 2, 1, False).state_dict()
    from oracle.ref_loader import load_reference_graphml, reference_available
    if reference_available():
        import inspect
        gml = load_reference_graphml()
        import magat_pathplanning_b200 as ours
        assert str(inspect.signature(ours.BatchLSIGF)) == str(inspect.signature(gml.BatchLSIGF))
        ref = gml.GraphFilterBatch(16, 24, 3, 1, True)
        assert list(ref.state_dict()) == list(m.state_dict()) and repr(ref) == repr(GraphFilterBatch(16, 24, 3, 1, True))


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["simt", "auto"])
@pytest.mark.parametrize("name", cases())
def test_cuda_matches_reference_golden(name, path):
    from magat_pathplanning_b200 import GraphFilterBatch
    d, meta = load(name)
    dev = torch.device("cuda:0")
    layer = GraphFilterBatch(meta["G"], meta["F"], meta["K"], 1, meta.get("bias", True))
    with torch.no_grad():
        layer.weight.copy_(d["param.weight"])
        if layer.bias is not None:
            layer.bias.copy_(d["param.bias"])
    layer = layer.to(dev)
    layer.path = path
    x = d["x"].to(dev).requires_grad_(True)
    layer.addGSO(d["S"].to(dev))
    y = layer(x)
    assert y.shape == d["y"].shape
    assert rel_err(y, d["y"]) < TOL
    y.backward(d["dy"].to(dev))
    assert rel_err(x.grad, d["grad.x"]) < TOL
    assert rel_err(layer.weight.grad, d["grad.weight"]) < TOL
    if layer.bias is not None:
        assert rel_err(layer.bias.grad, d["grad.bias"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("G,F,K,B,N", [(128, 128, 3, 6, 200), (128, 128, 2, 3, 1000), (64, 32, 4, 4, 77)])
def test_cuda_matches_oracle_larger(G, F, K, B, N):
    from magat_pathplanning_b200 import GraphFilterBatch
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(7 + N + K)
    S = orc.random_geometric_gso(B, N, generator=gen)
    x = torch.randn(B, N, G, generator=gen).permute(0, 2, 1)
    dy = torch.randn(B, F, N, generator=gen)
    torch.manual_seed(N)
    layer = GraphFilterBatch(G, F, K, 1, True)
    y_ref, g_ref = orc.lsigf_fwd_bwd(x, S, layer.weight.detach(), layer.bias.detach(), dy)
    layer = layer.to(dev)
    xd = x.to(dev).requires_grad_(True)
    layer.addGSO(S.to(dev))
    y = layer(xd)
    y.backward(dy.to(dev))
    assert rel_err(y, y_ref) < TOL
    assert rel_err(xd.grad, g_ref["x"]) < TOL
    assert rel_err(layer.weight.grad, g_ref["weight"]) < TOL
    assert rel_err(layer.bias.grad, g_ref["bias"]) < TOL
