"""CPU oracle for MAGAT's batched graph-attention layer.  TEST INFRASTRUCTURE ONLY.

This file is a checker, not a product path: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``magat_pathplanning_b200``) never imports anything under ``oracle/``.

It restates, densely and with the same torch ops (matmul / softmax / leaky_relu), the
algorithm of the reference (paths relative to the reference checkout):

* ``utils/graphUtils/graphML.py:1180-1286``  learnAttentionGSOBatch_KeyQuery
* ``utils/graphUtils/graphML.py:713-823``    learnAttentionGSOBatch (GAT_modified)
* ``utils/graphUtils/graphML.py:1724-1827``  graphAttentionLSIGFBatch_{KeyQuery,modified}
* ``utils/graphUtils/graphML.py:4636-4671``  GraphFilterBatchAttentional.forward

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports the unmodified reference
from /root/reference) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against them.

Gradients are obtained by torch autograd through this dense restatement, which is what
the reference itself does (``loss.backward()`` over the same ATen graph).
"""
from __future__ import annotations

import torch

ZERO_TOL = 1e-9      # graphML.py:45  zeroTolerance
BIG = 1e12           # graphML.py:46  infiniteNumber
LEAKY_SLOPE = 0.2    # graphML.py:713 negative_slope default, never overridden

MODES = ("KeyQuery", "GAT_modified")


def edge_mask(S: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[B,E,N,N] GSO -> [B,1,1,N,N] 0/1 mask; values of S are never used otherwise.

    graphML.py:1274-1276 / :808-809.  NaN compares false, hence "no edge".
    """
    m = S.detach().abs().sum(dim=1) > ZERO_TOL
    return m.to(dtype)[:, None, None]


def masked_row_softmax(e: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """softmax over the last axis restricted to mask, zero rows where nothing is allowed.

    graphML.py:1278-1286 / :811-823: scores are zeroed then pushed to -1e12 off the mask,
    soft-maxed over j and multiplied by the mask again (a fully masked row becomes the
    uniform 1/N and is then wiped to zero).
    """
    a = torch.softmax(e * mask - (1.0 - mask) * BIG, dim=-1)
    return a * mask


def attention_keyquery(x, weight, S):
    """e[b,p,i,j] = x_i^T W_p x_j (no LeakyReLU, no scaling).  graphML.py:1246-1266.

    x [B,G,N]; weight [P,E=1,G,G]; S [B,1,N,N] -> [B,P,1,N,N]
    """
    B, G, N = x.shape
    P = weight.shape[0]
    xq = x.reshape(B, 1, 1, G, N)
    xk = xq.transpose(3, 4)                                   # [B,1,1,N,G]
    Wx = torch.matmul(weight.reshape(1, P, 1, G, G), xq)      # [B,P,1,G,N]
    e = torch.matmul(xk, Wx)                                  # [B,P,1,N,N]
    return masked_row_softmax(e, edge_mask(S, x.dtype))


def attention_gat_modified(x, mixer, weight, weight_bias, S):
    """Additive GAT scores.  graphML.py:777-796.

    z = W_p x + wb_p (F x N);  s[i,j] = a2.z_i + a1.z_j with a1 = mixer[..., :F],
    a2 = mixer[..., F:];  e = LeakyReLU_0.2(s).
    """
    B, G, N = x.shape
    P, E, F, _ = weight.shape
    z = torch.matmul(weight.reshape(1, P, E, F, G), x.reshape(B, 1, 1, G, N))
    z = z + weight_bias.reshape(1, P, E, F, 1)
    a1 = mixer[:, :, :F].reshape(1, P, E, 1, F)
    a2 = mixer[:, :, F:].reshape(1, P, E, 1, F)
    col = torch.matmul(a1, z)                                 # [B,P,E,1,N]  (index j)
    row = torch.matmul(a2, z).transpose(3, 4)                 # [B,P,E,N,1]  (index i)
    e = torch.nn.functional.leaky_relu(col + row, negative_slope=LEAKY_SLOPE)
    return masked_row_softmax(e, edge_mask(S, x.dtype))


def lsigf_attention(filterWeight, x, aij, bias):
    """K-tap filter over the learned attention.  graphML.py:1745-1775.

    u_0 = x, u_k = u_{k-1} @ A  ([G,N] @ [N,N]: node j gathers from senders i with the
    row-normalised weight A[i,j]);  y[b,p,f,n] = sum_{k,g} h[p,f,0,k,g] u_k[b,p,g,n] + bias[f].
    """
    P, F, E, K, G = filterWeight.shape
    B, _, N = x.shape
    u = x.reshape(B, 1, 1, G, N).expand(B, P, E, G, N)
    taps = [u]
    for _ in range(1, K):
        u = torch.matmul(u, aij)
        taps.append(u)
    z = torch.stack(taps, dim=3)                              # [B,P,E,K,G,N]
    z = z.permute(0, 1, 5, 2, 3, 4).reshape(B, P, N, E * K * G)
    h = filterWeight.reshape(1, P, F, E * K * G).transpose(2, 3)
    y = torch.matmul(z, h).transpose(2, 3)                    # [B,P,F,N]
    if bias is not None:
        y = y + bias
    return y


def gat_layer_forward(x, S, params, *, mode="KeyQuery", concatenate=True, return_pre=False):
    """Whole-layer forward, GraphFilterBatchAttentional.forward (graphML.py:4636-4671).

    x [B,G,Nin]; S [B,1,N,N] (any float dtype); params: dict with mixer, weight_bias,
    filterWeight, bias (or None), weight.  Returns (y, aij) with y laid out exactly like
    the reference: concat -> [B,P*F,N] (a permuted view over [B,N,P*F] memory, channel
    p*F+f), mean -> [B,F,N]; ReLU applied as the reference does.
    """
    if mode not in MODES:
        raise ValueError(mode)
    B, _, Nin = x.shape
    N = S.shape[2]
    if Nin < N:                                               # graphML.py:4642-4646
        x = torch.cat((x, x.new_zeros(B, x.shape[1], N - Nin)), dim=2)
    if mode == "KeyQuery":
        aij = attention_keyquery(x, params["weight"], S)
    else:
        aij = attention_gat_modified(x, params["mixer"], params["weight"],
                                     params["weight_bias"], S)
    y = lsigf_attention(params["filterWeight"], x, aij, params.get("bias"))
    pre = y
    P, F = params["filterWeight"].shape[:2]
    if concatenate:                                           # graphML.py:4654-4662
        y = torch.relu(y)
        y = y.permute(0, 3, 1, 2).reshape(B, N, P * F).permute(0, 2, 1)
    else:                                                     # graphML.py:4665-4667
        y = torch.relu(y.mean(dim=1))
    if Nin < N:                                               # graphML.py:4669-4670
        y = y[:, :, :Nin]
    if return_pre:
        return y, aij, pre
    return y, aij


def gat_layer_fwd_bwd(x, S, params, dy, *, mode="KeyQuery", concatenate=True):
    """Forward + autograd backward.  Returns (y, aij, grads) with grads a dict holding
    'x' and one entry per parameter (None where the reference leaves grad=None, e.g.
    mixer / weight_bias in KeyQuery mode)."""
    x = x.detach().clone().requires_grad_(True)
    p = {k: (v.detach().clone().requires_grad_(True) if v is not None else None)
         for k, v in params.items()}
    y, aij = gat_layer_forward(x, S, p, mode=mode, concatenate=concatenate)
    y.backward(dy)
    grads = {"x": x.grad}
    for k, v in p.items():
        grads[k] = None if v is None else v.grad
    return y.detach(), aij.detach(), grads


def init_params(G, F, K, P, *, mode="KeyQuery", bias=True, E=1, generator=None,
                dtype=torch.float32, weight_bias_std=0.0):
    """Parameters with the reference's shapes and init (graphML.py:4579-4612):
    U(-s, s) with s = 1/sqrt(G*P); weight_bias zeros unless weight_bias_std > 0."""
    if mode == "KeyQuery" and F != G:
        raise ValueError("KeyQuery needs F == G (graphML.py:1728,1765)")
    s = 1.0 / (G * P) ** 0.5

    def U(*shape):
        return (torch.rand(*shape, generator=generator, dtype=torch.float64) * 2 - 1).mul(s).to(dtype)

    params = {
        "mixer": U(P, E, 2 * F),
        "weight_bias": torch.zeros(P, E, F, dtype=dtype),
        "filterWeight": U(P, F, E, K, G),
        "bias": U(F, 1) if bias else None,
        "weight": U(P, E, G, G) if mode == "KeyQuery" else U(P, E, F, G),
    }
    if weight_bias_std > 0:
        params["weight_bias"] = (torch.randn(P, E, F, generator=generator, dtype=torch.float64)
                                 * weight_bias_std).to(dtype)
    return params


def random_geometric_gso(B, N, *, comm_radius=7.0, density=0.025, width=None,
                         generator=None, dtype=torch.float32, normalize=True):
    """Synthetic GSO batch shaped like the simulator's (utils/new_simulator.py:816-846):
    N distinct integer cells on a w x w map, edge iff distance < comm_radius, zero diagonal,
    optionally divided by the largest eigenvalue.  Returns [B,1,N,N]."""
    if width is None:
        width = max(2, int(round((N / density) ** 0.5)))
    out = torch.zeros(B, 1, N, N, dtype=dtype)
    for b in range(B):
        cells = torch.randperm(width * width, generator=generator)[:N]
        pos = torch.stack((cells // width, cells % width), dim=1).to(torch.float64)
        d = torch.cdist(pos, pos)
        A = ((d < comm_radius) & ~torch.eye(N, dtype=torch.bool)).to(torch.float64)
        if normalize and A.sum() > 0:
            A = A / torch.linalg.eigvalsh(A).abs().max().clamp_min(1e-12)
        out[b, 0] = A.to(dtype)
    return out


def gso_from_positions(pos, comm_radius):
    """SURVEY 8f row f1 -- the simulator's adjacency from agent positions, utils/new_simulator.py:823-827:
    ``distances = squareform(pdist(pos, 'euclidean')); W = (distances < commR); W -= diag(diag(W))``.
    pos: [B,N,2]; returns the 0/1 mask [B,1,N,N] in fp64 (the normalisation of :829-839 only rescales it and the
    attention layer only tests |S| > 1e-9)."""
    pos = pos.to(torch.float64)
    diff = pos[:, :, None, :] - pos[:, None, :, :]
    d = diff.pow(2).sum(dim=-1).sqrt()                  # what pdist computes, pair by pair, in fp64
    N = pos.shape[1]
    A = (d < float(comm_radius)) & ~torch.eye(N, dtype=torch.bool)
    return A.to(torch.float64).unsqueeze(1)



# ---- the non-attentional graph filter ------------------------------------------------------------

def lsigf_forward(x, S, weight, bias=None):
    """BatchLSIGF, graphML.py:5485-5579, restated densely:  u_0 = x, u_k = u_{k-1} S (the values of S, cast to fp32,
    :5569-5571),  y[b,f,n] = sum_{k,g} h[f,0,k,g] u_k[b,g,n] + bias[f]  (:5573-5577).

    x [B,G,N]; S [B,1,N,N] (fp32 or fp64); weight [F,1,K,G]; bias [F,1] or None -> y [B,F,N].  Parity pinned by
    tests/golden/lsigf_golden.npz (outputs of the unmodified reference, tests/golden/make_golden_lsigf.py)."""
    F, E, K, G = weight.shape
    assert E == 1 and S.shape[1] == 1 and x.shape[1] == G
    Sf = S[:, 0].float()
    u = x
    taps = [u]
    for _ in range(1, K):
        u = torch.matmul(u, Sf)
        taps.append(u)
    z = torch.stack(taps, dim=1)                              # [B,K,G,N]
    y = torch.einsum("fkg,bkgn->bfn", weight[:, 0], z)
    if bias is not None:
        y = y + bias
    return y


def lsigf_fwd_bwd(x, S, weight, bias, dy):
    """Forward + autograd backward of ``lsigf_forward``: (y, {"x", "weight", "bias"} gradients)."""
    xg = x.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    b = None if bias is None else bias.detach().clone().requires_grad_(True)
    y = lsigf_forward(xg, S, w, b)
    y.backward(dy)
    return y.detach(), {"x": xg.grad, "weight": w.grad, "bias": None if b is None else b.grad}


# ---- the GAT_origin ablation ----------------------------------------------------------------------

def origin_layer_forward(x, S, params, *, concatenate=True):
    """GraphFilterBatchAttentional_Origin.forward (graphML.py:4284-4311) over graphAttentionLSIGFBatch_Origin
    (:1939-2005) and learnAttentionGSOBatch_origin (:964-1070), dense.

    params: mixer [P,1,2F], weight [P,1,F,G], filterWeight [1,K] (scalar taps), bias [F,1] or None.
    Scores s[i,j] = a2.z_i + a1.z_j with z = W x (no bias), LeakyReLU 0.2; edge mask from S + I (:1019); the filter of
    head p and tap k is h[k] times the reference's reshape of W (:1964-1969)."""
    mixer, W, h, bias = params["mixer"], params["weight"], params["filterWeight"], params.get("bias")
    B, G, Nin = x.shape
    N = S.shape[2]
    if Nin < N:
        x = torch.cat((x, torch.zeros(B, G, N - Nin, dtype=x.dtype)), dim=2)
    P, E, F, _ = W.shape
    K = h.shape[1]
    S1 = S.to(torch.float32) + torch.eye(N).reshape(1, 1, N, N)
    z = torch.einsum("pfg,bgn->bpfn", W[:, 0], x)
    a1, a2 = mixer[:, 0, :F], mixer[:, 0, F:]
    s = torch.einsum("pf,bpfn->bpn", a2, z)[:, :, :, None] + torch.einsum("pf,bpfn->bpn", a1, z)[:, :, None, :]
    e = torch.nn.functional.leaky_relu(s, 0.2)
    mask = edge_mask(S1, x.dtype).reshape(B, 1, N, N)
    aij = masked_row_softmax(e, mask)                                   # B x P x N x N
    filt = h.reshape(1, 1, E, K, 1) * W.permute(0, 3, 1, 2).reshape(P, F, E, 1, G)      # P x F x E x K x G
    u = x[:, None].expand(B, P, G, N)
    y = torch.einsum("pfg,bpgn->bpfn", filt[:, :, 0, 0], u)
    for k in range(1, K):
        u = torch.matmul(u, aij)
        y = y + torch.einsum("pfg,bpgn->bpfn", filt[:, :, 0, k], u)
    if bias is not None:
        y = y + bias
    if concatenate:
        out = torch.relu(y).permute(0, 3, 1, 2).reshape(B, N, P * F).permute(0, 2, 1)
    else:
        out = torch.relu(y.mean(dim=1))
    if Nin < N:
        out = out[:, :, :Nin]
    return out, aij.reshape(B, P, 1, N, N)
