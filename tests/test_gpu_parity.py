"""Parity of the CUDA path (through the C ABI) with the golden vectors of the unmodified reference
and with the CPU oracle.  Tolerance: 1e-4 max-norm relative (BASELINE.json north_star) on outputs,
attention and every gradient; the fp32 SIMT path is expected to sit near 1e-6."""
import pickle

import numpy as np
import pytest
import torch

from conftest import golden_case_names
from oracle import gat_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-4
PARAMS = ("mixer", "weight_bias", "filterWeight", "bias", "weight")


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make_layer(meta, d, dev, path="simt"):
    from magat_pathplanning_b200 import GraphFilterBatchAttentional
    layer = GraphFilterBatchAttentional(meta["G"], meta["F"], meta["K"], meta["P"], 1, meta.get("bias", True),
                                        concatenate=meta["concat"], attentionMode=meta["mode"])
    with torch.no_grad():
        for k in PARAMS:
            if ("param." + k) in d:
                getattr(layer, k).copy_(d["param." + k])
    layer.path = path
    return layer.to(dev)


@pytest.mark.parametrize("path", ["simt", "auto"])
@pytest.mark.parametrize("name", golden_case_names())
def test_golden_forward_backward(golden, name, path):
    d, meta = golden.case(name)
    dev = torch.device("cuda:0")
    layer = make_layer(meta, d, dev, path=path)
    x = d["x"].to(dev).requires_grad_(True)
    layer.addGSO(d["S"].to(dev))
    y = layer(x)
    assert y.shape == d["y"].shape
    assert rel_err(y, d["y"]) < TOL
    aij = layer.aij
    assert isinstance(aij, np.ndarray) and aij.shape == tuple(d["aij"].shape)
    assert np.abs(aij - d["aij"].numpy()).max() < TOL
    ret = layer.returnAttentionGSO()
    assert list(ret.shape) == meta["returnAttentionGSO_shape"]
    assert np.abs(ret - d["aij"].numpy().mean(axis=1)).max() < TOL
    y.backward(d["dy"].to(dev))
    assert rel_err(x.grad, d["grad.x"]) < TOL
    for k in PARAMS:
        p = getattr(layer, k)
        if p is None:
            continue
        if k in meta["none_grads"]:
            assert p.grad is None, f"{k} must keep grad=None like the reference"
        else:
            assert rel_err(p.grad, d["grad." + k]) < TOL, k


def test_output_layout_matches_reference(golden):
    dev = torch.device("cuda:0")
    for name in ("kq_concat_c2", "kq_mean_c1"):
        d, meta = golden.case(name)
        layer = make_layer(meta, d, dev)
        layer.addGSO(d["S"].to(dev))
        with torch.no_grad():
            y = layer(d["x"].to(dev))
        B, C, N = y.shape
        if meta["concat"]:      # permuted view over [B,N,P*F] memory (graphML.py:4656-4662)
            assert y.stride() == (N * C, 1, C)
        else:                   # contiguous [B,F,N] (graphML.py:4665-4667)
            assert y.is_contiguous()


def test_x_layouts_and_gso_dtype(golden):
    """x as the planners pass it (permuted view of [B,N,G]) and as a plain contiguous [B,G,N] tensor."""
    dev = torch.device("cuda:0")
    d, meta = golden.case("kq_concat_n10")
    layer = make_layer(meta, d, dev)
    layer.addGSO(d["S"].double().to(dev))               # simulator GSOs are float64 (new_simulator.py:317)
    with torch.no_grad():
        xa = d["x"].to(dev).contiguous()                # [B,G,N] contiguous
        xb = d["x"].permute(0, 2, 1).contiguous().to(dev).permute(0, 2, 1)   # view of [B,N,G]
        assert not xb.is_contiguous()
        ya, yb = layer(xa), layer(xb)
    assert rel_err(ya, d["y"]) < TOL and rel_err(yb, d["y"]) < TOL
    assert torch.equal(ya, yb)


def test_gso_scale_invariance_and_isolated_rows(golden):
    dev = torch.device("cuda:0")
    d, meta = golden.case("kq_concat_n10")
    layer = make_layer(meta, d, dev)
    x = d["x"].to(dev)
    with torch.no_grad():
        layer.addGSO(d["S"].to(dev))
        y1 = layer(x)
        a1 = layer.aij
        layer.addGSO((d["S"] * 3.7).to(dev))
        y2 = layer(x)
    assert torch.equal(y1, y2)
    mask = (d["S"].abs() > 1e-9)[:, 0].numpy()
    assert np.all(a1[:, :, 0][np.broadcast_to(~mask[:, None], a1[:, :, 0].shape)] == 0.0)
    iso = ~mask.any(-1)
    rows = a1[:, :, 0].sum(-1)
    assert np.all(rows[np.broadcast_to(iso[:, None], rows.shape)] == 0)


def test_requires_grad_gating(golden):
    dev = torch.device("cuda:0")
    d, meta = golden.case("gm_concat_fneg")
    layer = make_layer(meta, d, dev)
    layer.filterWeight.requires_grad_(False)
    layer.weight.requires_grad_(False)
    layer.addGSO(d["S"].to(dev))
    x = d["x"].to(dev)                                    # no grad on x either
    y = layer(x)
    y.backward(d["dy"].to(dev))
    assert layer.filterWeight.grad is None and layer.weight.grad is None
    assert rel_err(layer.mixer.grad, d["grad.mixer"]) < TOL
    assert rel_err(layer.bias.grad, d["grad.bias"]) < TOL
    assert rel_err(layer.weight_bias.grad, d["grad.weight_bias"]) < TOL


def test_functional_surface(golden):
    from magat_pathplanning_b200 import (graphAttentionLSIGFBatch_KeyQuery, graphAttentionLSIGFBatch_modified,
                                         learnAttentionGSOBatch, learnAttentionGSOBatch_KeyQuery)
    dev = torch.device("cuda:0")
    for name, fn, att_fn in (("kq_concat_n10", graphAttentionLSIGFBatch_KeyQuery, learnAttentionGSOBatch_KeyQuery),
                             ("gm_concat_fneg", graphAttentionLSIGFBatch_modified, learnAttentionGSOBatch)):
        d, meta = golden.case(name)
        p = {k: (d["param." + k].to(dev) if ("param." + k) in d else None) for k in PARAMS}
        x, S = d["x"].to(dev), d["S"].to(dev)
        y, aij = fn(p["filterWeight"], x, p["mixer"], p["weight"], p["weight_bias"], S, b=p["bias"])
        _, aij_ref, pre = orc.gat_layer_forward(d["x"], d["S"], {k: d.get("param." + k) for k in PARAMS},
                                                mode=meta["mode"], concatenate=True, return_pre=True)
        assert y.shape == pre.shape and rel_err(y, pre) < TOL
        assert aij.shape == aij_ref.shape and (aij.cpu() - aij_ref).abs().max() < TOL
        a2 = att_fn(x, p["mixer"], p["weight"], p["weight_bias"], S)      # the reference's argument order for both
        assert (a2.cpu() - aij_ref).abs().max() < TOL


def test_custom_nonlinearity_and_pickle(golden):
    dev = torch.device("cuda:0")
    d, meta = golden.case("kq_concat_n10")
    layer = make_layer(meta, d, dev)
    layer.nonlinearity = torch.tanh
    layer.addGSO(d["S"].to(dev))
    with torch.no_grad():
        y = layer(d["x"].to(dev))
    _, _, pre = orc.gat_layer_forward(d["x"], d["S"], {k: d.get("param." + k) for k in PARAMS},
                                      mode="KeyQuery", concatenate=True, return_pre=True)
    B, P, F, N = pre.shape
    want = torch.tanh(pre).permute(0, 3, 1, 2).reshape(B, N, P * F).permute(0, 2, 1)
    assert rel_err(y, want) < TOL
    # a nonlinearity that is NOT elementwise sees [B,P,F,N], as in the reference (graphML.py:4656): softmax over the
    # features of each head
    layer.nonlinearity = lambda t: torch.softmax(t, dim=2)
    with torch.no_grad():
        y = layer(d["x"].to(dev))
    want = torch.softmax(pre, dim=2).permute(0, 3, 1, 2).reshape(B, N, P * F).permute(0, 2, 1)
    assert list(y.stride()) == list(want.stride()) and rel_err(y, want) < TOL
    layer.nonlinearity = torch.nn.functional.relu
    clone = pickle.loads(pickle.dumps(layer))            # mp.spawn pickles the model (agents/...GAT.py:720-728)
    clone.addGSO(d["S"].to(dev))
    with torch.no_grad():
        assert rel_err(clone(d["x"].to(dev)), d["y"]) < TOL


def test_cpu_tensors_raise(golden):
    d, meta = golden.case("kq_concat_n10")
    layer = make_layer(meta, d, torch.device("cpu"))
    layer.addGSO(d["S"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(d["x"])


@pytest.mark.parametrize("mode,concat,G,F,K,P,B,N,gso", [
    ("KeyQuery", True, 128, 128, 3, 4, 8, 200, "geometric"),
    ("KeyQuery", False, 32, 32, 2, 4, 4, 100, "geometric"),
    ("GAT_modified", True, 128, 128, 3, 4, 4, 100, "geometric"),
    ("KeyQuery", True, 64, 64, 3, 2, 2, 70, "full"),
    ("GAT_modified", False, 24, 40, 4, 3, 2, 65, "full"),
])
def test_oracle_parity_larger(mode, concat, G, F, K, P, B, N, gso):
    """Sizes the oracle finishes in seconds, beyond the committed golden cases."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1337 + N + G)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    S = torch.ones(B, 1, N, N) if gso == "full" else orc.random_geometric_gso(B, N, generator=gen)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    dy = torch.randn(B, P * F if concat else F, N, generator=gen)
    y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(x, S, params, dy, mode=mode, concatenate=concat)
    meta = dict(G=G, F=F, K=K, P=P, concat=concat, mode=mode)
    layer = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev)
    xd = x.to(dev).requires_grad_(True)
    layer.addGSO(S.to(dev))
    y = layer(xd)
    y.backward(dy.to(dev))
    assert rel_err(y, y_ref) < TOL
    assert (torch.from_numpy(layer.aij) - aij_ref).abs().max() < TOL
    assert rel_err(xd.grad, g_ref["x"]) < TOL
    for k in PARAMS:
        if g_ref[k] is not None:
            assert rel_err(getattr(layer, k).grad, g_ref[k]) < TOL, k


# ---- tcgen05 (bf16 hi/lo split) projections ---------------------------------------------------
# Same 1e-4 bar; the three-pass split is expected near 1e-5.

@pytest.mark.parametrize("name", ["kq_mean_c1", "kq_concat_c2", "kq_n130_g128"])
def test_tc_path_golden(golden, name):
    d, meta = golden.case(name)
    dev = torch.device("cuda:0")
    layer = make_layer(meta, d, dev, path="tcgen05")
    x = d["x"].to(dev).requires_grad_(True)
    layer.addGSO(d["S"].to(dev))
    y = layer(x)
    assert rel_err(y, d["y"]) < TOL
    assert np.abs(layer.aij - d["aij"].numpy()).max() < TOL
    y.backward(d["dy"].to(dev))
    assert rel_err(x.grad, d["grad.x"]) < TOL
    for k in PARAMS:
        if ("grad." + k) in d:
            assert rel_err(getattr(layer, k).grad, d["grad." + k]) < TOL, k


@pytest.mark.parametrize("mode,concat,G,F,K,P,B,N", [
    ("KeyQuery", True, 128, 128, 3, 4, 8, 200),
    ("KeyQuery", False, 128, 128, 2, 2, 3, 150),
    ("GAT_modified", True, 128, 128, 3, 4, 4, 100),
    ("GAT_modified", False, 64, 128, 2, 3, 5, 77),
    ("KeyQuery", True, 256, 256, 2, 1, 2, 50),
    ("KeyQuery", True, 128, 128, 4, 2, 2, 60),          # K = 4: more tap blocks than fit TMEM next to the accumulators
    ("GAT_modified", True, 128, 128, 4, 4, 2, 40),
    ("KeyQuery", True, 128, 128, 3, 3, 2, 70),          # P = 3: none of the {1, 2, 4}-head sparse kernels
])
def test_tc_path_oracle(mode, concat, G, F, K, P, B, N):
    """Forward AND backward of the tcgen05 path (both attention modes) against the CPU oracle."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4242 + N + G)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    dy = torch.randn(B, P * F if concat else F, N, generator=gen)
    # no gradient into outputs next to the ReLU kink (a 5e-6 difference in y would flip relu'(y) there)
    _, _, pre = orc.gat_layer_forward(x, S, params, mode=mode, concatenate=concat, return_pre=True)
    pre_out = pre.reshape(B, P * F, N) if concat else pre.mean(dim=1)
    dy = dy * (pre_out.abs() > 1e-3)
    y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(x, S, params, dy, mode=mode, concatenate=concat)
    meta = dict(G=G, F=F, K=K, P=P, concat=concat, mode=mode)
    layer = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path="tcgen05")
    layer.addGSO(S.to(dev))
    xd = x.to(dev).requires_grad_(True)
    y = layer(xd)
    y.backward(dy.to(dev))
    with torch.no_grad():
        layer.path = "simt"
        y_simt = layer(x.to(dev))
    e_tc, e_simt = rel_err(y, y_ref), rel_err(y_simt, y_ref)
    print(f"tcgen05 err {e_tc:.2e}  simt err {e_simt:.2e}")
    assert e_tc < TOL
    assert (torch.from_numpy(layer.aij) - aij_ref).abs().max() < TOL
    assert rel_err(xd.grad, g_ref["x"]) < TOL
    for k in PARAMS:
        if g_ref[k] is not None:
            assert rel_err(getattr(layer, k).grad, g_ref[k]) < TOL, k


@pytest.mark.parametrize("path", ["auto", "fused"])
def test_oracle_at_the_graded_shape(path):
    """B x N = 160 x 1000, G = F = 128, K = 3, P = 4, concat, KeyQuery -- the north-star shape, persistent CTAs wrapping
    their rings many times -- against the ORACLE: instances are independent, so the oracle runs on three of them (its
    dense [P,N,N] temporaries fit) and, with dy zero everywhere else, pins y, aij, dx AND every parameter gradient of
    the whole at-scale run."""
    from magat_pathplanning_b200.graphML import Adjacency, attention_dense
    from bench import synth_gso
    dev = torch.device("cuda:0")
    G = F = 128
    B, N, K, P = 160, 1000, 3, 4
    gen = torch.Generator().manual_seed(2024)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen, weight_bias_std=0.1)
    S = synth_gso(B, N, 200, dev, torch.Generator(device=dev).manual_seed(17))
    x_mem = torch.relu(torch.randn(B, N, G, generator=gen))
    pick = [0, 77, 159]
    xs, Ss = x_mem[pick].permute(0, 2, 1), S[pick].cpu()
    dys = torch.randn(len(pick), P * F, N, generator=gen)
    _, _, pre = orc.gat_layer_forward(xs, Ss, params, mode="KeyQuery", concatenate=True, return_pre=True)
    dys = dys * (pre.reshape(len(pick), P * F, N).abs() > 1e-3)
    y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(xs, Ss, params, dys, mode="KeyQuery", concatenate=True)
    dy = torch.zeros(B, P * F, N)
    dy[pick] = dys
    layer = make_layer(dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery"),
                       {"param." + k: v for k, v in params.items() if v is not None}, dev, path=path)
    xd = x_mem.to(dev).permute(0, 2, 1).requires_grad_(True)
    layer.addGSO(S)
    y = layer(xd)
    y.backward(dy.to(dev))
    assert rel_err(y[pick], y_ref) < TOL
    last = layer._last
    idx = torch.tensor(pick, device=dev)
    adj = last.adj
    sub = Adjacency(len(pick), N, adj.D, adj.nbr_out[idx].contiguous(), adj.nbr_in[idx].contiguous(),
                    adj.slot_in[idx].contiguous(), None)
    aij = attention_dense(last.att[idx].contiguous(), sub).cpu()
    assert (aij - aij_ref).abs().max() < TOL
    assert rel_err(xd.grad[pick], g_ref["x"]) < TOL
    rest = torch.ones(B, dtype=torch.bool)
    rest[pick] = False
    assert float(xd.grad[rest.to(dev)].abs().max()) == 0.0
    for k in PARAMS:
        if g_ref[k] is not None:
            assert rel_err(getattr(layer, k).grad, g_ref[k]) < TOL, k


@pytest.mark.parametrize("B,N,K,P", [(160, 1000, 3, 4), (130, 777, 2, 2)])
def test_tc_path_matches_simt_at_scale(B, N, K, P):
    """The tcgen05 pipelines (TMA rings, converter groups, persistent CTAs) against the fp32 SIMT kernels of the
    same library at a size where every CTA wraps its rings many times (>= 10^5 node rows; the oracle's dense
    [B,P,N,N] temporaries do not fit there).  The SIMT path itself is pinned to the oracle / golden vectors above.
    A ring-parity bug in the weight-gradient kernel only showed from ~10^5 rows up."""
    dev = torch.device("cuda:0")
    G = F = 128
    gen = torch.Generator().manual_seed(99 + N)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen, weight_bias_std=0.1)
    from bench import synth_gso
    gdev = torch.Generator(device=dev).manual_seed(7 + N)
    S = synth_gso(B, N, int(round((N / 0.025) ** 0.5)), dev, gdev)
    x_mem = torch.relu(torch.randn(B, N, G, generator=gen)).to(dev)
    dy_mem = torch.randn(B, N, P * F, generator=gen).to(dev)
    meta = dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery")
    # The two paths differ by ~5e-6 in y, which flips relu'(y) for the few hundred outputs that sit that close to
    # zero; their gradient would then differ by a whole dy entry.  No gradient flows into outputs below 1e-3.
    probe = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path="simt")
    probe.addGSO(S)
    with torch.no_grad():
        y0 = probe(x_mem.permute(0, 2, 1))
    dy_mem = dy_mem * (y0.permute(0, 2, 1) > 1e-3)
    del probe, y0
    out = {}
    for path in ("simt", "tcgen05"):
        layer = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path=path)
        for rep in range(2):                     # twice back to back: kernels of consecutive steps overlap at the seams
            for p_ in layer.parameters():
                p_.grad = None
            xd = x_mem.permute(0, 2, 1).detach().requires_grad_(True)
            layer.addGSO(S)
            y = layer(xd)
            y.backward(dy_mem.permute(0, 2, 1))
        torch.cuda.synchronize()
        out[path] = dict(y=y.detach(), dx=xd.grad, dH=layer.filterWeight.grad, dW=layer.weight.grad, db=layer.bias.grad)
    errs = {k: rel_err(out["tcgen05"][k], out["simt"][k]) for k in out["simt"]}
    print("tcgen05 vs simt: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert all(v < TOL for v in errs.values()), errs


@pytest.mark.parametrize("B,N,P", [(3, 50, 2), (7, 333, 4), (2, 64, 1)])
def test_relu_bit_mask_of_the_projection_epilogue(B, N, P):
    """Training forward on the tcgen05 path: the K-tap projection leaves (y > 0) as one bit per element for the
    backward (word [m >> 5][c], bit m & 31), bit-exact against the y it stored, also where the row count is no multiple
    of 32 or 64; the backward that reads the bits gives the gradients of the one that reads y (SIMT path)."""
    dev = torch.device("cuda:0")
    G = F = 128
    K = 3
    gen = torch.Generator().manual_seed(31 + N)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen).to(dev)
    x = torch.randn(B, G, N, generator=gen).to(dev)
    meta = dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery")
    layer = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path="auto")
    layer.addGSO(S)
    xd = x.clone().requires_grad_(True)
    y = layer(xd)
    node, seen = y.grad_fn, 0
    while node is not None and "_GATFunction" not in type(node).__name__ and seen < 8:
        node, seen = node.next_functions[0][0], seen + 1
    bits = node.saved_tensors[-1]
    assert bits is not None and bits.dtype == torch.int32
    rows, C = B * N, P * F
    words = bits.view(-1, C)[: (rows + 31) // 32].cpu().numpy().astype(np.uint32)
    got = ((words[:, None, :] >> np.arange(32, dtype=np.uint32)[None, :, None]) & 1).reshape(-1, C)
    want = (y.detach().permute(0, 2, 1).reshape(rows, C) > 0).cpu().numpy()
    assert np.array_equal(got[:rows].astype(bool), want)
    assert not got[rows:].any()                       # bits past the last row are zero
    # gradients: bits (tcgen05 backward) against y (SIMT backward); dy kept away from the ReLU kink
    dy = torch.randn(B, C, N, generator=gen).to(dev) * (y.detach() > 1e-3)
    y.backward(dy)
    ref = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path="simt")
    ref.addGSO(S)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(dy)
    assert rel_err(xd.grad, xr.grad) < TOL
    for k in PARAMS:
        if getattr(ref, k).grad is not None:
            assert rel_err(getattr(layer, k).grad, getattr(ref, k).grad) < TOL, k


def test_tc_path_rejects_uncovered_shape(golden):
    from magat_pathplanning_b200._cabi import MagatError
    d, meta = golden.case("kq_concat_n10")          # G = 16
    dev = torch.device("cuda:0")
    layer = make_layer(meta, d, dev, path="tcgen05")
    layer.addGSO(d["S"].to(dev))
    with pytest.raises(MagatError, match="not covered"):
        layer(d["x"].to(dev))


@pytest.mark.parametrize("N", [1, 5, 12, 36, 100, 130, 1000, 1004])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gso_scan_and_neighbour_lists(N, dtype):
    """GSO -> bit masks -> padded lists against the dense mask |S| > 1e-9 (graphML.py:1274-1276), incl. NaN,
    negative and sub-tolerance entries, asymmetric masks, every scan kernel variant (scalar, vector, TMA ring)."""
    from magat_pathplanning_b200 import build_adjacency
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(99 + N)
    B = 3
    S = (torch.rand(B, 1, N, N, generator=gen) < min(0.5, 6.0 / max(N, 1))).to(dtype)
    S = S * torch.randn(B, 1, N, N, generator=gen).to(dtype)
    if N > 4:
        S[:, :, 2, 3] = float("nan")
        S[:, :, 3, 4] = 5e-10
        S[:, :, 4, 1] = -3e-9
        S[:, :, 0, :] = 0
    mask = (S.abs() > 1e-9)[:, 0]                               # [B,N,N]
    adj = build_adjacency(S.to(dev), with_slot_out=True)
    out, inn, slot, sout = adj.nbr_out.cpu(), adj.nbr_in.cpu(), adj.slot_in.cpu(), adj.slot_out.cpu()
    D = adj.D
    assert D % 4 == 0 and D >= max(int(mask.sum(2).max()), int(mask.sum(1).max()), 1)
    for b in range(B):
        for i in range(N):
            want = torch.nonzero(mask[b, i]).flatten().tolist()
            got = [v for v in out[b, i].tolist() if v >= 0]
            assert got == want and out[b, i, len(want):].eq(-1).all()
            want_in = torch.nonzero(mask[b, :, i]).flatten().tolist()
            got_in = [v for v in inn[b, i].tolist() if v >= 0]
            assert got_in == want_in
            for s, src in enumerate(want_in):
                assert out[b, src, slot[b, i, s]] == i
            for s, dst in enumerate(want):
                assert inn[b, dst, sout[b, i, s]] == i
        if N >= 1000:
            break                                               # one instance is enough at this size


# ---- single-launch small-graph inference kernel (the simulator's B = 1, N <= 64 use) --------------------

@pytest.mark.parametrize("N,width,R", [(1, 4, 7.0), (10, 20, 7.0), (37, 30, 7.0), (100, 50, 7.0), (1000, 200, 7.0),
                                       (130, 40, 3.5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_adjacency_from_positions(N, width, R, dtype):
    """lists built on the device from positions == lists of the dense GSO the simulator formula
    gives (oracle.gso_from_positions, pinned to scipy's pdist in the CPU suite)."""
    from magat_pathplanning_b200 import build_adjacency, build_adjacency_from_positions
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(N * 7 + 1)
    B = 5
    pos = torch.empty(B, N, 2, dtype=torch.float64)
    for b in range(B):
        cells = torch.randperm(width * width, generator=gen)[:N]
        pos[b] = torch.stack((cells // width, cells % width), dim=1).to(torch.float64)
    if N == 130:
        pos += torch.rand(pos.shape, generator=gen, dtype=torch.float64)       # off-lattice
    pos = pos.to(dtype)
    S = orc.gso_from_positions(pos, R)
    a = build_adjacency(S.to(dev))
    p_ = build_adjacency_from_positions(pos.to(dev), R)
    assert a.D == p_.D
    for name in ("nbr_out", "nbr_in", "slot_in"):
        assert torch.equal(getattr(a, name), getattr(p_, name)), name


@pytest.mark.parametrize("case", ["exact_radius_lattice", "huge_radius", "tiny_radius_many_cells", "nan_and_inf",
                                  "negative_coordinates", "zero_radius", "n3000"])
def test_adjacency_from_positions_edge_cases(case):
    """The cell-list builder against the dense formula where cells could go wrong: pairs at exactly the radius (lattice
    positions: d = 7 is NOT an edge, d just below is), one cell for everything, a bounding box of more cells than the
    kernel bins (it then takes all pairs), non-finite positions (no edge, as the comparison is false), negative and large
    coordinates, a radius of zero, and the largest N the kernel takes."""
    from magat_pathplanning_b200 import build_adjacency, build_adjacency_from_positions
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(len(case))
    B, N, R = 3, 200, 7.0
    pos = torch.rand(B, N, 2, generator=gen, dtype=torch.float64) * 60
    if case == "exact_radius_lattice":
        pos = torch.floor(pos)                                   # integer coordinates: many pairs at distance exactly 7, 5-5-7.07...
        pos[:, 1] = pos[:, 0] + torch.tensor([7.0, 0.0])
        pos[:, 2] = pos[:, 0] + torch.tensor([0.0, -7.0])
        pos[:, 3] = pos[:, 0] + torch.tensor([7.0 - 1e-12, 0.0])
    elif case == "huge_radius":
        R = 1e6
    elif case == "tiny_radius_many_cells":
        R, pos = 0.05, pos * 10                                  # 600 / 0.05 = 12000 cells per side
        pos[:, 1] = pos[:, 0] + 0.03
    elif case == "nan_and_inf":
        pos[0, 5, 0] = float("nan")
        pos[1, 7, 1] = float("inf")
        pos[2, 9] = float("-inf")
    elif case == "negative_coordinates":
        pos = pos - 1e5
        pos[:, 1] = pos[:, 0] + 6.999
    elif case == "zero_radius":
        R = 0.0
    elif case == "n3000":
        B, N = 2, 3000
        pos = torch.rand(B, N, 2, generator=gen, dtype=torch.float64) * 340
    S = orc.gso_from_positions(pos, R)
    a = build_adjacency(S.to(dev))
    p_ = build_adjacency_from_positions(pos.to(dev), R)
    assert a.D == p_.D
    for name in ("nbr_out", "nbr_in", "slot_in"):
        assert torch.equal(getattr(a, name), getattr(p_, name)), name


def test_layer_from_positions_matches_dense_gso(golden):
    from magat_pathplanning_b200 import GraphFilterBatchAttentional
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(11)
    B, N, G, K, P = 3, 90, 128, 3, 4
    cells = torch.stack([torch.randperm(50 * 50, generator=gen)[:N] for _ in range(B)])
    pos = torch.stack((cells // 50, cells % 50), dim=2).to(torch.float32)
    S = orc.gso_from_positions(pos, 7.0).to(torch.float32)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1).to(dev)
    torch.manual_seed(3)
    layer = GraphFilterBatchAttentional(G, G, K, P, 1, True, concatenate=True, attentionMode="KeyQuery").to(dev)
    with torch.no_grad():
        layer.addGSO(S.to(dev))
        y_dense = layer(x)
        layer.addGSOFromPositions(pos.to(dev), 7.0)
        y_pos = layer(x)
        att = layer.returnAttentionGSO()
    assert torch.equal(y_dense, y_pos)
    assert att.shape == (B, 1, N, N)


SMALL_CASES = [n for n in golden_case_names() if n not in ("kq_b32p4_n100", "kq_n130_g128", "gm_n70_g64")]


@pytest.mark.parametrize("name", SMALL_CASES)
def test_small_graph_kernel_golden(golden, name):
    from magat_pathplanning_b200 import _cabi
    d, meta = golden.case(name)
    dev = torch.device("cuda:0")
    layer = make_layer(meta, d, dev, path="auto")
    layer.addGSO(d["S"].to(dev))
    L = _cabi.lib()
    with torch.no_grad():
        c0 = L.magat_launch_count()
        y = layer(d["x"].to(dev))
        launches = L.magat_launch_count() - c0
    assert launches == 1, "inference on a small graph must be ONE kernel launch"
    assert y.shape == d["y"].shape and rel_err(y, d["y"]) < TOL
    if meta["concat"] and "Nin" not in meta:
        B, C, N = y.shape
        assert y.stride() == (N * C, 1, C) or N == 1
    assert np.abs(layer.aij - d["aij"].numpy()).max() < TOL
    assert np.abs(layer.returnAttentionGSO() - d["aij"].numpy().mean(axis=1)).max() < TOL


def test_small_graph_kernel_not_used_when_training(golden):
    from magat_pathplanning_b200 import _cabi
    d, meta = golden.case("kq_concat_n10")
    dev = torch.device("cuda:0")
    layer = make_layer(meta, d, dev, path="auto")
    layer.addGSO(d["S"].to(dev))
    L = _cabi.lib()
    c0 = L.magat_launch_count()
    y = layer(d["x"].to(dev))                       # parameters require grad -> autograd path
    assert L.magat_launch_count() - c0 > 1
    y.backward(d["dy"].to(dev))
    assert rel_err(layer.filterWeight.grad, d["grad.filterWeight"]) < TOL


@pytest.mark.parametrize("N,promise", [(10, None), (100, 24)])
def test_layer_is_cuda_graph_capturable(N, promise):
    """No host synchronisation in addGSO / forward / backward when the list width is known up front (N <= 32, or a
    max_degree promise): the whole training step replays as one CUDA graph and reproduces the eager result."""
    dev = torch.device("cuda:0")
    G = F = 128
    K, P, B = 3, 4, 6
    gen = torch.Generator().manual_seed(40 + N)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen).to(dev)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1).to(dev)
    dy = torch.randn(B, P * F, N, generator=gen).to(dev)
    layer = make_layer(dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery"),
                       {"param." + k: v for k, v in params.items() if v is not None}, dev, path="auto")
    layer.max_degree = promise
    out = {}

    def step():
        for p_ in layer.parameters():
            p_.grad = None
        xg = x.detach().requires_grad_(True)
        layer.addGSO(S)
        y = layer(xg)
        y.backward(dy)
        out["y"], out["dx"] = y.detach(), xg.grad
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in out.items()}
    eager["dH"] = layer.filterWeight.grad.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["y"], eager["y"]) and torch.equal(out["dx"], eager["dx"])
    assert rel_err(layer.filterWeight.grad, eager["dH"]) < 1e-6
    layer._last.adj.check_degree()                      # the promise held


def test_gso_in_host_memory_gives_the_same_adjacency_and_output(golden):
    """addGSO with a CPU tensor: the mask is packed on the host (magat_gso_pack_host) and transposed on the device, or --
    with few host cores per process -- the dense tensor is copied and scanned on the device; same lists either way."""
    from magat_pathplanning_b200 import build_adjacency, build_adjacency_host
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(12)
    for N, dtype in ((1000, torch.float32), (77, torch.float64), (10, torch.float32)):
        S = orc.random_geometric_gso(3, N, generator=gen).to(dtype)
        S[0, 0, 1, 2] = float("nan")
        S[1, 0, 2, 1] = -3e-9
        S[2, 0, 0, 1] = 5e-10
        a = build_adjacency(S.to(dev))
        # threads = 2: packed on the host whatever the box; -1: the dense-copy route (few cores per process); 0: policy
        for threads in (2, -1, 0):
            b = build_adjacency_host(S, dev, threads=threads)
            assert a.D == b.D
            for k in ("nbr_out", "nbr_in", "slot_in"):
                assert torch.equal(getattr(a, k), getattr(b, k)), (N, k, threads)
    d, meta = golden.case("kq_concat_c2")
    layer = make_layer(meta, d, dev)
    layer.addGSO(d["S"])                                   # stays on the CPU
    x = d["x"].to(dev).requires_grad_(True)
    y = layer(x)
    assert rel_err(y, d["y"]) < TOL
    y.backward(d["dy"].to(dev))
    assert rel_err(x.grad, d["grad.x"]) < TOL


@pytest.mark.parametrize("mode,B,N,K,P,A", [("KeyQuery", 5, 200, 3, 4, 5), ("GAT_modified", 3, 77, 2, 2, 5),
                                           ("KeyQuery", 2, 64, 3, 1, 8), ("KeyQuery", 7, 333, 1, 4, 3)])
def test_layer_with_fused_action_head(mode, B, N, K, P, A):
    """layer + the planner's linear action head in one pass (y never written) against the ORACLE's
    layer output pushed through the same nn.Linear on the CPU, and against the two-step route of the CUDA layer; the
    decoded actions are argmax softmax (utils/new_simulator.py:863-869)."""
    dev = torch.device("cuda:0")
    G = F = 128
    gen = torch.Generator().manual_seed(77 + N + A)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    head = torch.nn.Linear(P * F, A)
    with torch.no_grad():
        head.weight.copy_(torch.randn(A, P * F, generator=gen) * 0.05)
        head.bias.copy_(torch.randn(A, generator=gen) * 0.1)
    y_ref, _ = orc.gat_layer_forward(x, S, params, mode=mode, concatenate=True)
    with torch.no_grad():
        logits_ref = head(y_ref.permute(0, 2, 1).reshape(B * N, -1))
    meta = dict(G=G, F=F, K=K, P=P, concat=True, mode=mode)
    layer = make_layer(meta, {"param." + k: v for k, v in params.items() if v is not None}, dev, path="auto")
    layer.addGSO(S.to(dev))
    head_d = torch.nn.Sequential(torch.nn.Linear(P * F, A)).to(dev)
    head_d[0].load_state_dict(head.state_dict())
    with torch.no_grad():
        logits, actions = layer.forward_actions(x.to(dev), head_d, return_actions=True)
        y2 = layer(x.to(dev))
        logits2 = head_d(y2.permute(0, 2, 1).reshape(B * N, -1))
    assert logits.shape == (B * N, A) and actions.shape == (B * N,) and actions.dtype == torch.int32
    assert rel_err(logits, logits_ref) < TOL
    assert rel_err(logits, logits2) < TOL
    assert torch.equal(actions.long(), torch.max(logits, 1)[1])
    # the attention of the call is kept like forward() keeps it
    assert (torch.from_numpy(layer.aij) - orc.gat_layer_forward(x, S, params, mode=mode, concatenate=True)[1]).abs().max() < TOL
    # with autograd on (or any shape the fused head does not take) the two-step route answers
    xg = x.to(dev).requires_grad_(True)
    lg = layer.forward_actions(xg, head_d)
    assert lg.requires_grad and rel_err(lg, logits_ref) < TOL


def test_fused_action_head_at_the_graded_shape():
    """B x N = 160 x 1000 (persistent CTAs wrap their rings many times; the last tile is ragged in neither 32 nor 64):
    the fused head against the ORACLE on two instances and against the two-step route on all of them."""
    from bench import synth_gso
    dev = torch.device("cuda:0")
    G = F = 128
    B, N, K, P, A = 160, 1000, 3, 4, 5
    gen = torch.Generator().manual_seed(909)
    params = orc.init_params(G, F, K, P, mode="KeyQuery", generator=gen, weight_bias_std=0.1)
    S = synth_gso(B, N, 200, dev, torch.Generator(device=dev).manual_seed(23))
    x_mem = torch.relu(torch.randn(B, N, G, generator=gen))
    head = torch.nn.Linear(P * F, A)
    with torch.no_grad():
        head.weight.copy_(torch.randn(A, P * F, generator=gen) * 0.05)
    layer = make_layer(dict(G=G, F=F, K=K, P=P, concat=True, mode="KeyQuery"),
                       {"param." + k: v for k, v in params.items() if v is not None}, dev, path="auto")
    layer.addGSO(S)
    head_d = torch.nn.Linear(P * F, A).to(dev)
    head_d.load_state_dict(head.state_dict())
    xd = x_mem.to(dev).permute(0, 2, 1)
    with torch.no_grad():
        logits, actions = layer.forward_actions(xd, head_d, return_actions=True)
        logits2 = head_d(layer(xd).permute(0, 2, 1).reshape(B * N, -1))
    assert rel_err(logits, logits2) < 2e-5
    pick = [0, 159]
    y_ref, _ = orc.gat_layer_forward(x_mem[pick].permute(0, 2, 1), S[pick].cpu(), params, mode="KeyQuery", concatenate=True)
    with torch.no_grad():
        ref = head(y_ref.permute(0, 2, 1).reshape(len(pick) * N, -1))
    got = logits.view(B, N, A)[pick].reshape(-1, A)
    assert rel_err(got, ref) < TOL
    # decode: equal to the argmax of the reference logits wherever the top two are not within rounding of each other
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref.abs().max()
    assert torch.equal(actions.view(B, N)[pick].reshape(-1).cpu().long()[clear], ref.argmax(1)[clear])


@pytest.mark.parametrize("mode,K,P", [("KeyQuery", 3, 4), ("KeyQuery", 2, 1), ("GAT_modified", 2, 4)])
def test_head_mean_at_the_tensor_core_shapes(mode, K, P):
    """Heads averaged (the reference's CLI default) with G = F = 128: routed through the concat path's tcgen05 kernels
    plus elementwise mean / ReLU.  Forward, output layout (contiguous [B,F,N], graphML.py:4665-4667) and every gradient
    against the oracle."""
    dev = torch.device("cuda:0")
    G = F = 128
    B, N = 8, 300
    gen = torch.Generator().manual_seed(555 + K + P)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    dy = torch.randn(B, F, N, generator=gen)
    _, _, pre = orc.gat_layer_forward(x, S, params, mode=mode, concatenate=False, return_pre=True)
    dy = dy * (pre.mean(dim=1).abs() > 1e-3)
    y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(x, S, params, dy, mode=mode, concatenate=False)
    layer = make_layer(dict(G=G, F=F, K=K, P=P, concat=False, mode=mode),
                       {"param." + k: v for k, v in params.items() if v is not None}, dev, path="auto")
    layer.addGSO(S.to(dev))
    xd = x.to(dev).requires_grad_(True)
    y = layer(xd)
    assert y.shape == (B, F, N) and y.is_contiguous()
    assert rel_err(y, y_ref) < TOL
    assert (torch.from_numpy(layer.aij) - aij_ref).abs().max() < TOL
    y.backward(dy.to(dev))
    assert rel_err(xd.grad, g_ref["x"]) < TOL
    for k in PARAMS:
        if g_ref[k] is not None:
            assert rel_err(getattr(layer, k).grad, g_ref[k]) < TOL, k
        else:
            assert getattr(layer, k).grad is None, k


@pytest.mark.parametrize("mode,concat,G,F,K,P", [("KeyQuery", False, 32, 32, 2, 4), ("KeyQuery", True, 64, 64, 3, 4),
                                                 ("GAT_modified", True, 32, 64, 3, 2), ("GAT_modified", False, 48, 128, 2, 1),
                                                 ("KeyQuery", True, 16, 16, 1, 3)])
def test_narrow_layers_zero_padded_onto_the_128_feature_kernels(mode, concat, G, F, K, P):
    """From 32768 node rows on, layers with fewer than 128 features run zero padded on the 128-feature kernels
    (graphML.py::_padded_layer): forward, layout, attention and every gradient against the oracle -- padding must not
    show anywhere."""
    from magat_pathplanning_b200 import graphML as ours
    dev = torch.device("cuda:0")
    B, N = 4, 150
    gen = torch.Generator().manual_seed(321 + G + F)
    params = orc.init_params(G, F, K, P, mode=mode, generator=gen, weight_bias_std=0.1)
    S = orc.random_geometric_gso(B, N, generator=gen)
    x = torch.relu(torch.randn(B, N, G, generator=gen)).permute(0, 2, 1)
    _, _, pre = orc.gat_layer_forward(x, S, params, mode=mode, concatenate=concat, return_pre=True)
    pre_out = pre.reshape(B, P * F, N) if concat else pre.mean(dim=1)
    dy = torch.randn(pre_out.shape, generator=gen) * (pre_out.abs() > 1e-3)
    y_ref, aij_ref, g_ref = orc.gat_layer_fwd_bwd(x, S, params, dy, mode=mode, concatenate=concat)
    layer = make_layer(dict(G=G, F=F, K=K, P=P, concat=concat, mode=mode),
                       {"param." + k: v for k, v in params.items() if v is not None}, dev, path="auto")
    old = ours._PAD_MIN_ROWS
    ours._PAD_MIN_ROWS = 1                       # (the test batch is small: force the route)
    try:
        layer.addGSO(S.to(dev))
        xd = x.to(dev).requires_grad_(True)
        y = layer(xd)
        assert y.shape == y_ref.shape and list(y.stride()) == list(y_ref.stride())
        assert rel_err(y, y_ref) < TOL
        assert (torch.from_numpy(layer.aij) - aij_ref).abs().max() < TOL
        y.backward(dy.to(dev))
    finally:
        ours._PAD_MIN_ROWS = old
    assert rel_err(xd.grad, g_ref["x"]) < TOL
    gmax = max(float(v.abs().max()) for v in g_ref.values() if v is not None)
    for k in PARAMS:
        if g_ref[k] is None:
            assert getattr(layer, k).grad is None, k
        elif float(g_ref[k].abs().max()) < 1e-6 * gmax:
            assert float(getattr(layer, k).grad.abs().max()) < 1e-5 * gmax, k
        else:
            assert tuple(getattr(layer, k).grad.shape) == tuple(g_ref[k].shape)
            assert rel_err(getattr(layer, k).grad, g_ref[k]) < TOL, k
