"""The single-launch fused forward (csrc/gat_fused.cu, magat_gat_forward_fused) against the CPU oracle and against
the multi-launch path of the same library.  Tolerance 1e-4 max-norm relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import gat_oracle as orc
from test_gpu_parity import PARAMS, TOL, make_layer, rel_err

pytestmark = pytest.mark.gpu


def _layer(params, meta, dev, path):
    return make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path=path)


@pytest.mark.parametrize("mode,K,P,B,N", [
    ("KeyQuery", 3, 4, 21, 200),          # 21 instances over 18 teams: three teams reuse their scratch
    ("KeyQuery", 2, 2, 5, 132),
    ("KeyQuery", 1, 1, 3, 64),
    ("KeyQuery", 3, 1, 2, 1000),
    ("GAT_modified", 3, 4, 6, 200),
    ("GAT_modified", 2, 2, 40, 68),
])
def test_fused_vs_oracle(mode, K, P, B, N):
    dev = torch.device("cuda:0")
    G = F = 128
    gen = torch.Generator().manual_seed(31 + N + K)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    dy = torch.randn(B, P * F, N, generator=gen)
    nb = min(B, 4)                                    # the oracle's dense temporaries: a few instances are enough
    dy[nb:] = 0
    # no gradient into outputs next to the ReLU kink: a 5e-6 difference in y would flip relu'(y) there
    _, _, pre = orc.gat_layer_forward(x[:nb], S[:nb], params, mode=mode, concatenate=True, return_pre=True)
    dy[:nb] *= (pre.reshape(nb, P * F, N).abs() > 1e-3)
    y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(x[:nb], S[:nb], params, dy[:nb], mode=mode, concatenate=True)
    meta = dict(G=G, F=F, K=K, P=P, concat=True, mode=mode)
    layer = _layer(params, meta, dev, "fused")
    xd = x.to(dev).requires_grad_(True)
    layer.addGSO(S.to(dev))
    y = layer(xd)
    y.backward(dy.to(dev))
    assert rel_err(y[:nb], y_ref) < TOL
    assert (torch.from_numpy(layer.aij[:nb]) - aij_ref).abs().max() < TOL
    assert rel_err(xd.grad[:nb], g_ref["x"]) < TOL
    assert float(xd.grad[nb:].abs().max()) == 0.0 if B > nb else True
    for k in PARAMS:
        if g_ref[k] is not None:
            assert rel_err(getattr(layer, k).grad, g_ref[k]) < TOL, k
    # inference mode (nothing saved, R and taps stay in the team scratch) gives the same output
    with torch.no_grad():
        layer.addGSO(S.to(dev))
        y2 = layer(x.to(dev))
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("mode", ["KeyQuery", "GAT_modified"])
def test_fused_matches_multi_launch_path(mode):
    """Same library, two routes: everything the fused launch leaves behind (lists, attention, taps, R, y, gradients)
    against the scan / lists / projection / attention / gather / projection launches."""
    dev = torch.device("cuda:0")
    G = F = 128
    K, P, B, N = 3, 4, 40, 1000
    gen = torch.Generator().manual_seed(5)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    from bench import synth_gso
    S = synth_gso(B, N, 200, dev, torch.Generator(device=dev).manual_seed(11))
    x_mem = torch.relu(torch.randn(B, N, G, generator=gen)).to(dev)
    dy_mem = torch.randn(B, N, P * F, generator=gen).to(dev)
    meta = dict(G=G, F=F, K=K, P=P, concat=True, mode=mode)
    probe = _layer(params, meta, dev, "simt")
    probe.addGSO(S)
    with torch.no_grad():
        y0 = probe(x_mem.permute(0, 2, 1))
    dy_mem = dy_mem * (y0.permute(0, 2, 1) > 1e-3)        # no gradient into outputs next to the ReLU kink
    out = {}
    for path in ("tcgen05", "fused"):
        layer = _layer(params, meta, dev, path)
        xd = x_mem.permute(0, 2, 1).detach().requires_grad_(True)
        layer.addGSO(S)
        y = layer(xd)
        y.backward(dy_mem.permute(0, 2, 1))
        torch.cuda.synchronize()
        out[path] = dict(y=y.detach(), dx=xd.grad, dH=layer.filterWeight.grad, dW=layer.weight.grad,
                         db=layer.bias.grad, att=layer._last.att, adj=layer._last.adj)
    a, b = out["tcgen05"]["adj"], out["fused"]["adj"]
    D = min(a.D, b.D)
    for k in ("nbr_out", "nbr_in", "slot_in"):
        la, lb = getattr(a, k), getattr(b, k)
        assert torch.equal(la[..., :D], lb[..., :D]), k
        assert bool((la[..., D:] <= 0).all()) and bool((lb[..., D:] <= 0).all())
    assert float((out["tcgen05"]["att"][:, :, :D] - out["fused"]["att"][:, :, :D]).abs().max()) < 1e-5
    errs = {k: rel_err(out["fused"][k], out["tcgen05"][k]) for k in ("y", "dx", "dH", "dW", "db")}
    print("fused vs multi-launch: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert all(v < TOL for v in errs.values()), errs


def test_fused_f64_gso_and_degree_retry():
    dev = torch.device("cuda:0")
    G = F = 128
    K, P, B, N = 2, 4, 4, 128
    gen = torch.Generator().manual_seed(77)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen)
    # a ring lattice with 24 neighbours per agent: over the first degree cap (16), under the second (32)
    idx = torch.arange(N)
    d = (idx[:, None] - idx[None, :]).abs()
    d = torch.minimum(d, N - d)
    S = ((d > 0) & (d <= 12)).double()[None, None].repeat(B, 1, 1, 1) * 0.05
    S[1, 0, 5, 6] = float("nan")                       # NaN is "no edge" (graphML.py:1274)
    S[2, 0, 7, 8] = -0.3                               # negative weights are edges
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    y_ref, aij_ref = orc.gat_layer_forward(x, torch.nan_to_num(S.float(), nan=0.0), params, mode="KeyQuery",
                                           concatenate=True)
    layer = _layer(params, dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery"), dev, "fused")
    layer.addGSO(S.to(dev))
    with torch.no_grad():
        y = layer(x.to(dev))
    assert layer._last.adj.D == 24
    assert rel_err(y, y_ref) < TOL
    assert (torch.from_numpy(layer.aij) - aij_ref).abs().max() < TOL
    # promised degree bound: no read-back, same result
    layer.max_degree = 24
    layer.addGSO(S.to(dev))
    with torch.no_grad():
        assert torch.equal(layer(x.to(dev)), y)


def test_fused_falls_back_when_a_vertex_has_more_than_32_neighbours():
    dev = torch.device("cuda:0")
    G = F = 128
    K, P, B, N = 2, 2, 2, 64
    gen = torch.Generator().manual_seed(78)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen)
    S = torch.ones(B, 1, N, N)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    y_ref, _ = orc.gat_layer_forward(x, S, params, mode="KeyQuery", concatenate=True)
    layer = _layer(params, dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery"), dev, "fused")
    layer.addGSO(S.to(dev))
    xd = x.to(dev).requires_grad_(True)
    y = layer(xd)
    assert layer._last.adj.D == 64
    assert rel_err(y, y_ref) < TOL
    y.sum().backward()
    assert xd.grad is not None


def test_fused_teams_of_sixteen():
    """fused_team = 16: twice the CTAs per instance, half the instances in flight (the scratch then fits L2)."""
    dev = torch.device("cuda:0")
    G = F = 128
    K, P, B, N = 3, 4, 11, 300
    gen = torch.Generator().manual_seed(61)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen).to(dev)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1).to(dev)
    layer = _layer(params, dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery"), dev, "fused")
    with torch.no_grad():
        layer.addGSO(S)
        y8 = layer(x)
        layer.fused_team = 16
        layer.addGSO(S)
        y16 = layer(x)
    assert torch.equal(y8, y16)
    y_ref, _ = orc.gat_layer_forward(x[:3].cpu(), S[:3].cpu(), params, mode="KeyQuery", concatenate=True)
    assert rel_err(y16[:3], y_ref) < TOL


def test_fused_is_one_launch():
    from magat_pathplanning_b200 import _cabi
    dev = torch.device("cuda:0")
    G = F = 128
    gen = torch.Generator().manual_seed(3)
    params = orc.init_params(G, F, 3, 4, mode="KeyQuery", generator=gen)
    S = orc.random_geometric_gso(6, 256, generator=gen).to(dev)
    x = torch.relu(torch.randn(6, 256, G, generator=gen)).permute(0, 2, 1).to(dev)
    layer = _layer(params, dict(G=G, F=F, K=3, P=4, concat=True, mode="KeyQuery"), dev, "fused")
    layer.max_degree = 16                  # promised bound: not even the degree read-back remains
    L = _cabi.lib()
    with torch.no_grad():
        layer.addGSO(S)
        layer(x)
        c0 = L.magat_launch_count()
        layer.addGSO(S)
        layer(x)
        assert L.magat_launch_count() - c0 == 1
