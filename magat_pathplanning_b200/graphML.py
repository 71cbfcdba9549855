"""Host-side mirror of the reference's operator interface for the batched graph-attention path.

Same names, argument meaning and error behaviour as ``utils/graphUtils/graphML.py`` of
proroklab/magat_pathplanning for:

* ``GraphFilterBatchAttentional``            (graphML.py:4506-4685)  nn.Module: params, addGSO, forward, returnAttentionGSO
* ``graphAttentionLSIGFBatch_KeyQuery``      (graphML.py:1724-1775)
* ``graphAttentionLSIGFBatch_modified``      (graphML.py:1777-1827)
* ``learnAttentionGSOBatch_KeyQuery``        (graphML.py:1180-1286)
* ``learnAttentionGSOBatch``                 (graphML.py:713-823)

All arithmetic happens in ``lib/libmagat_gat.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/magat_gat.h``); torch is used for device memory, streams and autograd bookkeeping only.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi

zeroTolerance = 1e-9     # graphML.py:45
infiniteNumber = 1e12    # graphML.py:46

#: no host sync for the ELL width up to this node count (width = N: no list can overflow, and the lane-per-slot
#: kernels still apply)
_NOSYNC_N = 32

_PATH = {"auto": _cabi.PATH_AUTO, "simt": _cabi.PATH_SIMT, "tcgen05": _cabi.PATH_TCGEN05, "fused": _cabi.PATH_AUTO}

#: path="fused" takes the single-launch fused forward (csrc/gat_fused.cu).  "auto" only does so when this switch is on
#: (MAGAT_FUSED_AUTO=1) and the graphs have at least _FUSED_MIN_N agents: on B200 the multi-launch path is currently the
#: faster of the two, so it stays the default
_FUSED_AUTO = os.environ.get("MAGAT_FUSED_AUTO", "0") not in ("", "0")
_FUSED_MIN_N = 128
#: degree cap D the fused kernel is first tried with (its lists are [B][N][D]); a graph that exceeds it is redone
#: with 32 and then through the general path
_FUSED_D0 = 16

_ws_cache = {}


def _workspace(dev, nbytes):
    """Scratch of the fused forward, one per (device, stream): reused call after call so it stays L2 resident."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes + 1024:
        t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        _ws_cache[key] = t
    off = (-t.data_ptr()) % 1024
    return t[off:off + nbytes]


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"magat_pathplanning_b200: {what} must be a CUDA tensor (sm_100a); "
                           "this package has no CPU path")


def _check_params(x, filterWeight, mixer, weight, weight_bias, bias, need_cuda=True):
    """Every parameter must be fp32 on x's device: the kernels get raw device pointers (the reference would raise a
    dtype / device error from its first matmul)."""
    for name, t in (("filterWeight", filterWeight), ("mixer", mixer), ("weight", weight),
                    ("weight_bias", weight_bias), ("bias", bias)):
        if t is None:
            continue
        if need_cuda:
            _require_cuda(t, name)
        if t.device != x.device:
            raise RuntimeError(f"magat_pathplanning_b200: {name} is on {t.device}, x on {x.device}")
        if t.dtype != torch.float32:
            raise RuntimeError(f"magat_pathplanning_b200: {name} is {t.dtype}; the layer computes in float32 "
                               "(call .float() on the module)")


class Adjacency:
    """Neighbour lists of one GSO batch (device tensors; layouts in include/magat_gat.h)."""

    __slots__ = ("B", "N", "D", "nbr_out", "nbr_in", "slot_in", "slot_out", "stats")

    def __init__(self, B, N, D, nbr_out, nbr_in, slot_in, slot_out, stats=None):
        self.B, self.N, self.D = B, N, D
        self.nbr_out, self.nbr_in, self.slot_in, self.slot_out = nbr_out, nbr_in, slot_in, slot_out
        self.stats = stats          # device int32 {max out-degree, max in-degree, edges, symmetric} of the scan, or None

    def record_stream(self, stream):
        """The lists were built on another stream (e.g. by a worker thread preparing the next chunk): tell the caching
        allocator that ``stream`` uses them too."""
        for t in (self.nbr_out, self.nbr_in, self.slot_in, self.slot_out, self.stats):
            if t is not None:
                t.record_stream(stream)
        return self

    def check_degree(self):
        """Synchronising check of a ``max_degree`` promise: raises when a neighbour list did not fit its D slots."""
        if self.stats is not None:
            h = self.stats.cpu()
            if max(int(h[0]), int(h[1])) > self.D:
                raise RuntimeError(f"max_degree promise broken: a node has {max(int(h[0]), int(h[1]))} neighbours, "
                                   f"the lists hold {self.D}")


_stats0 = {}


def _stats_init(dev):
    """{0, 0, 0, 1}: the scan's statistics block before the call (one device-to-device copy of a cached constant, so
    the layer stays capturable in a CUDA graph)."""
    t = _stats0.get(dev.index)
    if t is None:
        t = torch.tensor([0, 0, 0, 1], dtype=torch.int32, device=dev)
        _stats0[dev.index] = t
    return t.clone()


def build_adjacency(S: torch.Tensor, max_degree: Optional[int] = None, nonzero: bool = False,
                    with_slot_out: bool = False, self_loops: bool = False) -> Adjacency:
    """[B,1,N,N] dense GSO -> neighbour lists.  Only ``|S| > 1e-9`` matters (graphML.py:1274-1276):
    NaN is "no edge", negative weights are edges.  S is read once, by one kernel.

    The list width D comes from N (N <= 32), from the caller's ``max_degree`` promise, or -- the only case with a
    host synchronisation -- from the degree statistics of the scan."""
    _require_cuda(S, "the GSO")
    assert len(S.shape) == 4
    B, E, N = S.shape[0], S.shape[1], S.shape[2]
    assert S.shape[3] == N
    if E != 1:
        raise NotImplementedError("edge_features E != 1 is not supported (the planners use E = 1, "
                                  "graphs/models/decentralplanner_GAT.py:179)")
    S = S.detach()
    if S.dtype not in (torch.float32, torch.float64):
        S = S.to(torch.float32)
    if not S.is_contiguous():
        S = S.contiguous()
    dev = S.device
    L = _cabi.lib()
    W = (N + 31) // 32
    with torch.cuda.device(dev):
        st = _stream(dev)
        rowbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
        colbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
        stats = _stats_init(dev)
        scan = L.magat_gso_scan_nonzero if nonzero else L.magat_gso_scan    # nonzero: BatchLSIGF's "S itself" predicate
        dt = _cabi.DT_F32 if S.dtype == torch.float32 else _cabi.DT_F64
        _cabi.check(scan(S.data_ptr(), dt, B, N, rowbits.data_ptr(), colbits.data_ptr(), stats.data_ptr(), st))
        if self_loops:          # GAT_origin: the edges of S + I (graphML.py:1019)
            _cabi.check(L.magat_gso_self_loops(S.data_ptr(), dt, B, N, rowbits.data_ptr(), colbits.data_ptr(),
                                               stats.data_ptr(), st))
        return _lists_from_masks(rowbits, colbits, stats, B, N, dev, st, max_degree, with_slot_out)


def pack_gso_host(S: torch.Tensor, threads: int = 0) -> torch.Tensor:
    """Edge mask of a GSO that lives in HOST memory, packed on the host cores: [B,1,N,N] fp32 / fp64 CPU tensor ->
    pinned int32 [B,N,ceil(N/32)] (bit j % 32 of word j / 32 of row i set iff |S[i,j]| > 1e-9).  One streaming pass
    (magat_gso_pack_host: threads + AVX2); 32x fewer bytes then cross PCIe than with the dense fp32 operator."""
    assert len(S.shape) == 4 and S.shape[1] == 1 and S.shape[2] == S.shape[3]
    if S.is_cuda:
        raise RuntimeError("pack_gso_host: the GSO is already on the device; use build_adjacency")
    S = S.detach()
    if S.dtype not in (torch.float32, torch.float64):
        S = S.to(torch.float32)
    if not S.is_contiguous():
        S = S.contiguous()
    B, N = S.shape[0], S.shape[2]
    pin = torch.cuda.is_available()
    bits = torch.empty((B, N, (N + 31) // 32), dtype=torch.int32, pin_memory=pin)
    _cabi.check(_cabi.lib().magat_gso_pack_host(S.data_ptr(), _cabi.DT_F32 if S.dtype == torch.float32 else _cabi.DT_F64,
                                                B * N, N, bits.data_ptr(), int(threads)))
    return bits


def build_adjacency_from_rowbits(bits: torch.Tensor, device, max_degree: Optional[int] = None) -> Adjacency:
    """Neighbour lists from a packed row mask [B,N,W] (``pack_gso_host``; host or device): H2D copy of the mask on the
    current stream, transposition into the column mask and the usual list build on the device."""
    dev = torch.device(device)
    B, N, W = bits.shape
    assert W == (N + 31) // 32 and bits.dtype == torch.int32
    L = _cabi.lib()
    with torch.cuda.device(dev):
        st = _stream(dev)
        rowbits = bits if bits.is_cuda else bits.to(dev, non_blocking=True)
        rowbits = rowbits.contiguous()
        colbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
        stats = _stats_init(dev)
        _cabi.check(L.magat_gso_from_rowbits(rowbits.data_ptr(), B, N, colbits.data_ptr(), stats.data_ptr(), st))
        return _lists_from_masks(rowbits, colbits, stats, B, N, dev, st, max_degree)


#: host cores a process needs to itself for the host-side packing to beat the copy engine (both read the same host
#: memory, the copy engine needs no core; 16-core bench host: packing wins with 1-2 processes, the dense copy with 4-8)
_PACK_MIN_CORES = 6


def host_cores_per_rank() -> int:
    """Cores this process may use, shared evenly with the other local ranks of a torchrun launch."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))


def build_adjacency_host(S: torch.Tensor, device, max_degree: Optional[int] = None, threads: int = 0) -> Adjacency:
    """``build_adjacency`` for a GSO in host memory.  With enough host cores per process the edge mask is packed on the
    host (N^2 / 8 bytes over PCIe); otherwise (or with ``threads < 0``) the dense tensor is copied and scanned on the
    device, which costs no core."""
    if threads < 0 or (threads == 0 and host_cores_per_rank() < _PACK_MIN_CORES):
        dev = torch.device(device)
        Sd = S.detach()
        if Sd.dtype not in (torch.float32, torch.float64):
            Sd = Sd.to(torch.float32)
        with torch.cuda.device(dev):
            return build_adjacency(Sd.to(dev, non_blocking=True), max_degree)
    return build_adjacency_from_rowbits(pack_gso_host(S, threads), device, max_degree)


def _lists_from_masks(rowbits, colbits, stats, B, N, dev, st, max_degree=None, with_slot_out=False):
    L = _cabi.lib()
    if N <= _NOSYNC_N:
        D = N
    elif max_degree:
        D = max(1, min(int(max_degree), N))
    else:
        h = stats.cpu()
        D = max(int(h[0]), int(h[1]), 1)
    D = (D + 3) // 4 * 4             # 16 B aligned neighbour rows (int4 / float4 loads in the kernels)
    nbr_out = torch.empty((B, N, D), dtype=torch.int32, device=dev)
    nbr_in = torch.empty((B, N, D), dtype=torch.int32, device=dev)
    slot_in = torch.empty((B, N, D), dtype=torch.int32, device=dev)
    # slot_out (position of a sender in its receivers' in-lists) only fed the receiver-major attention copy of the
    # round-1 kernels; nothing on the current path reads it, so it is built on request only
    slot_out = torch.empty((B, N, D), dtype=torch.int32, device=dev) if with_slot_out else None
    _cabi.check(L.magat_gso_build_ell(rowbits.data_ptr(), colbits.data_ptr(), B, N, D, nbr_out.data_ptr(),
                                      nbr_in.data_ptr(), slot_in.data_ptr(), _p(slot_out), st))
    return Adjacency(B, N, D, nbr_out, nbr_in, slot_in, slot_out, stats)


def build_adjacency_from_positions(pos: torch.Tensor, comm_radius: float, max_degree: Optional[int] = None) -> Adjacency:
    """neighbour lists straight from agent positions [B,N,2] (fp32 / fp64, CUDA), edge iff
    the Euclidean distance is < comm_radius, no self loops -- the mask utils/new_simulator.py:823-827 builds on the
    CPU before shipping a dense N x N GSO.  Equal to ``build_adjacency`` of that GSO."""
    _require_cuda(pos, "the positions")
    assert len(pos.shape) == 3 and pos.shape[2] == 2
    pos = pos.detach()
    if pos.dtype not in (torch.float32, torch.float64):
        pos = pos.to(torch.float64)
    if not pos.is_contiguous():
        pos = pos.contiguous()
    B, N = pos.shape[0], pos.shape[1]
    dev = pos.device
    L = _cabi.lib()
    W = (N + 31) // 32
    with torch.cuda.device(dev):
        st = _stream(dev)
        rowbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
        colbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
        stats = _stats_init(dev)
        _cabi.check(L.magat_gso_from_positions(pos.data_ptr(),
                                               _cabi.DT_F32 if pos.dtype == torch.float32 else _cabi.DT_F64, B, N,
                                               float(comm_radius), rowbits.data_ptr(), colbits.data_ptr(),
                                               stats.data_ptr(), st))
        return _lists_from_masks(rowbits, colbits, stats, B, N, dev, st, max_degree)


def _node_major(x: torch.Tensor):
    """x is logically [B,G,N]; the kernels want node-major rows.  The planners hand over a permuted
    view of contiguous [B,N,G] memory (decentralplanner_GAT.py:304) which is used in place."""
    xt = x.permute(0, 2, 1)
    G = xt.shape[2]
    if xt.shape[2] > 1 and xt.stride(2) != 1:
        xt = xt.contiguous()
    elif xt.stride(1) < G or (xt.shape[0] > 1 and xt.stride(0) < 0):
        xt = xt.contiguous()
    return xt


def _sn(xt):      # size-1 dims carry arbitrary strides
    return xt.stride(1) if xt.shape[1] > 1 else xt.shape[2]


def _sb(xt):
    return xt.stride(0) if xt.shape[0] > 1 else xt.shape[1] * _sn(xt)


class _Meta:
    __slots__ = ("mode", "concat", "relu", "path", "G", "F", "K", "P", "has_bias", "needs_grad")


class _FusedGSO:
    """The dense GSO on its way into the fused forward (which builds the neighbour lists itself); ``adj`` holds the
    lists the kernel wrote."""
    __slots__ = ("S", "D", "trusted", "adj", "overflow", "team")

    def __init__(self, S, D, trusted, team=0):
        self.S, self.D, self.trusted, self.adj, self.overflow, self.team = S, D, trusted, None, False, team


class _GATFunction(torch.autograd.Function):
    """x[B,G,N] -> y (concat: [B,P*F,N] view over [B,N,P*F]; mean: [B,F,N]) plus the sparse attention."""

    @staticmethod
    def forward(ctx, x, weight, mixer, weight_bias, filterWeight, bias, adj, meta: _Meta):
        if isinstance(adj, _FusedGSO):
            return _GATFunction._forward_fused(ctx, x, weight, mixer, weight_bias, filterWeight, bias, adj, meta)
        L = _cabi.lib()
        dev = x.device
        B, G, N = x.shape
        F, K, P, D = meta.F, meta.K, meta.P, adj.D
        xt = _node_major(x.detach())
        weight_c = weight.detach().contiguous()
        filt_c = filterWeight.detach().contiguous()
        mixer_c = None if mixer is None else mixer.detach().contiguous()
        wb_c = None if weight_bias is None else weight_bias.detach().contiguous()
        bias_c = None if bias is None else bias.detach().contiguous()
        C_out = P * F if meta.concat else F
        with torch.cuda.device(dev):
            if meta.concat:
                y_mem = torch.empty((B, N, C_out), dtype=torch.float32, device=dev)
                y = y_mem.permute(0, 2, 1)
            else:
                y_mem = torch.empty((B, C_out, N), dtype=torch.float32, device=dev)
                y = y_mem
            att = torch.empty((B, N, D, P), dtype=torch.float32, device=dev)
            ain = None          # (receiver-major attention copy: only the round-1 gather kernels used it)
            taps = torch.empty((B, N, P, max(K - 1, 1), G), dtype=torch.float32, device=dev) if K > 1 else None
            wprep = torch.empty(L.magat_gat_wprep_floats(G, F, K, P, meta.mode), dtype=torch.float32, device=dev)
            sproj = torch.empty((B, N, P, G if meta.mode == _cabi.MODE_KEYQUERY else 2), dtype=torch.float32,
                                device=dev)
            # training: the projection epilogue also leaves the ReLU mask as one bit per element, so that the dense
            # backward kernels need not read y again
            bits = None
            if meta.needs_grad and meta.relu and meta.concat:
                bits = torch.empty(L.magat_gat_relu_bits_words(B, N, P, F), dtype=torch.int32, device=dev)
            a = _cabi.FwdArgs(B=B, N=N, G=G, F=F, K=K, P=P, D=D, mode=meta.mode, concat=int(meta.concat),
                              relu=int(meta.relu), path=meta.path, reserved=0, relu_bits=_p(bits),
                              x=xt.data_ptr(), x_sb=_sb(xt), x_sn=_sn(xt),
                              nbr_out=adj.nbr_out.data_ptr(), nbr_in=adj.nbr_in.data_ptr(),
                              slot_in=adj.slot_in.data_ptr(), slot_out=_p(adj.slot_out),
                              weight=weight_c.data_ptr(), mixer=_p(mixer_c), weight_bias=_p(wb_c),
                              filterWeight=filt_c.data_ptr(), bias=_p(bias_c),
                              y=y_mem.data_ptr(), y_sb=y.stride(0), y_sn=y.stride(2), y_sc=y.stride(1),
                              att=att.data_ptr(), ain=_p(ain), taps=_p(taps), wprep=wprep.data_ptr(),
                              sproj=sproj.data_ptr())
            _cabi.check(L.magat_gat_forward(a, _stream(dev)))
            ctx.taps_valid = L.magat_gat_forward_taps_valid(a)
            if bits is not None and not L.magat_gat_forward_relu_bits_valid(a):
                bits = None
        ctx.meta, ctx.adj = meta, adj
        ctx.save_for_backward(xt, weight_c, mixer_c, wb_c, filt_c, y, att, taps, wprep, sproj, bits)
        ctx.mark_non_differentiable(att)
        ctx.set_materialize_grads(False)       # no zero-filled gradient tensor for the attention output
        return y, att

    @staticmethod
    def _forward_fused(ctx, x, weight, mixer, weight_bias, filterWeight, bias, gso, meta):
        """ONE launch from the dense GSO to y (magat_gat_forward_fused); leaves for backward exactly what the
        multi-launch forward leaves."""
        L = _cabi.lib()
        dev = x.device
        B, G, N = x.shape
        F, K, P = meta.F, meta.K, meta.P
        S = gso.S
        xt = _node_major(x.detach())
        weight_c = weight.detach().contiguous()
        filt_c = filterWeight.detach().contiguous()
        mixer_c = None if mixer is None else mixer.detach().contiguous()
        wb_c = None if weight_bias is None else weight_bias.detach().contiguous()
        bias_c = None if bias is None else bias.detach().contiguous()
        save = bool(meta.needs_grad)
        kq = meta.mode == _cabi.MODE_KEYQUERY
        with torch.cuda.device(dev):
            y_mem = torch.empty((B, N, P * F), dtype=torch.float32, device=dev)
            y = y_mem.permute(0, 2, 1)
            taps = sproj = wprep = None
            if save:
                taps = torch.empty((B, N, P, max(K - 1, 1), G), dtype=torch.float32, device=dev) if K > 1 else None
                sproj = torch.empty((B, N, P, G if kq else 2), dtype=torch.float32, device=dev)
                wprep = torch.empty(L.magat_gat_wprep_floats(G, F, K, P, meta.mode), dtype=torch.float32, device=dev)
            D = gso.D
            while True:
                lists = torch.empty((3, B, N, D), dtype=torch.int32, device=dev)
                att = torch.empty((B, N, D, P), dtype=torch.float32, device=dev)
                nbytes = L.magat_gat_fused_workspace_bytes(B, N, K, P, D, meta.mode, int(save), gso.team)
                ws = _workspace(dev, nbytes)
                a = _cabi.FusedArgs(B=B, N=N, G=G, F=F, K=K, P=P, D=D, mode=meta.mode, concat=1, relu=int(meta.relu),
                                    s_dtype=_cabi.DT_F32 if S.dtype == torch.float32 else _cabi.DT_F64, save=int(save),
                                    team=gso.team, reserved=0,
                                    S=S.data_ptr(), x=xt.data_ptr(), x_sb=_sb(xt), x_sn=_sn(xt),
                                    weight=weight_c.data_ptr(), mixer=_p(mixer_c), weight_bias=_p(wb_c),
                                    filterWeight=filt_c.data_ptr(), bias=_p(bias_c),
                                    y=y_mem.data_ptr(), y_sb=y.stride(0), y_sn=y.stride(2), y_sc=y.stride(1),
                                    nbr_out=lists[0].data_ptr(), nbr_in=lists[1].data_ptr(),
                                    slot_in=lists[2].data_ptr(), slot_out=None,
                                    att=att.data_ptr(), taps=_p(taps), sproj=_p(sproj), wprep=_p(wprep),
                                    workspace=ws.data_ptr(), ws_bytes=nbytes)
                _cabi.check(L.magat_gat_forward_fused(a, _stream(dev)))
                if D >= N or gso.trusted:
                    break                        # no list can be longer than D: nothing to check
                st = ws[:16].view(torch.int32).tolist()          # {max out-degree, max in-degree, overflows, watchdog}
                if st[2] == 0:
                    break
                need = (max(st[0], st[1]) + 3) // 4 * 4
                if need > 32:
                    gso.overflow = True          # a vertex with more than 32 neighbours: general path
                    gso.adj = build_adjacency(S)
                    return _GATFunction.forward(ctx, x, weight, mixer, weight_bias, filterWeight, bias, gso.adj, meta)
                D = need
        adj = Adjacency(B, N, D, lists[0], lists[1], lists[2], None)      # (nothing downstream needs slot_out)
        gso.adj = adj
        ctx.taps_valid = K - 1
        ctx.meta, ctx.adj = meta, adj
        ctx.save_for_backward(xt, weight_c, mixer_c, wb_c, filt_c, y, att, taps, wprep, sproj, None)
        ctx.mark_non_differentiable(att)
        ctx.set_materialize_grads(False)       # no zero-filled gradient tensor for the attention output
        return y, att

    @staticmethod
    def backward(ctx, dy, _datt):
        if dy is None:                                   # y took no part in the loss
            return (None,) * 8
        L = _cabi.lib()
        meta, adj = ctx.meta, ctx.adj
        xt, weight_c, mixer_c, wb_c, filt_c, y, att, taps, wprep, sproj, bits = ctx.saved_tensors
        dev = xt.device
        B, N, G = xt.shape
        F, K, P, D = meta.F, meta.K, meta.P, adj.D
        gm = meta.mode == _cabi.MODE_GAT_MODIFIED
        need = ctx.needs_input_grad
        # K == 1: the attention never reaches y, so the reference leaves weight/mixer/weight_bias grads None
        need_dx, need_dw = need[0], (need[1] and K > 1)
        need_dmix = gm and K > 1 and (need[2] or need[3])
        need_df, need_db = need[4], (need[5] and meta.has_bias)
        if dy.dtype != torch.float32:
            dy = dy.float()
        with torch.cuda.device(dev):
            def buf(*shape):
                return torch.empty(shape, dtype=torch.float32, device=dev)
            dx = buf(B, N, G) if need_dx else None
            dw = torch.empty_like(weight_c) if need_dw else None
            dmix = torch.empty_like(mixer_c) if need_dmix else None
            dwb = torch.empty_like(wb_c) if need_dmix else None
            df = torch.empty_like(filt_c) if need_df else None
            db = buf(F, 1) if need_db else None
            nws = L.magat_gat_backward_workspace_bytes(B, N, G, F, K, P, D, meta.mode)
            ws = torch.empty(nws + 256, dtype=torch.uint8, device=dev)         # ONE scratch allocation (gz, datt, rc, partial)
            ws = ws[(-ws.data_ptr()) % 256:]
            a = _cabi.BwdArgs(B=B, N=N, G=G, F=F, K=K, P=P, D=D, mode=meta.mode, concat=int(meta.concat),
                              relu=int(meta.relu), path=meta.path,
                              need_dx=int(need_dx), need_dweight=int(need_dw), need_dfilter=int(need_df),
                              need_dbias=int(need_db), need_dmixer=int(need_dmix),
                              taps_valid=int(ctx.taps_valid), reserved=0,
                              x=xt.data_ptr(), x_sb=_sb(xt), x_sn=_sn(xt),
                              nbr_out=adj.nbr_out.data_ptr(), nbr_in=adj.nbr_in.data_ptr(),
                              slot_in=adj.slot_in.data_ptr(),
                              weight=weight_c.data_ptr(), mixer=_p(mixer_c), weight_bias=_p(wb_c),
                              filterWeight=filt_c.data_ptr(),
                              y=y.data_ptr(), y_sb=y.stride(0), y_sn=y.stride(2), y_sc=y.stride(1),
                              att=att.data_ptr(), taps=_p(taps), wprep=wprep.data_ptr(), sproj=sproj.data_ptr(),
                              dy=dy.data_ptr(), dy_sb=dy.stride(0), dy_sn=dy.stride(2), dy_sc=dy.stride(1),
                              dx=_p(dx), dweight=_p(dw), dmixer=_p(dmix), dweight_bias=_p(dwb),
                              dfilterWeight=_p(df), dbias=_p(db),
                              gz=None, datt=None, rc=None, partial=None, relu_bits=_p(bits))
            _cabi.check(L.magat_gat_backward_ws(a, ws.data_ptr(), nws, _stream(dev)))
        gx = dx.permute(0, 2, 1) if need_dx else None
        # KeyQuery never touches mixer / weight_bias: the reference leaves their grad = None
        return (gx, dw, dmix if (need_dmix and need[2]) else None, dwb if (need_dmix and need[3]) else None,
                df, db, None, None)


class _HeadMean(torch.autograd.Function):
    """[B, P*F, N] (the concat layout: a view over [B,N,P*F] memory) -> act(mean over the heads) as the reference's
    contiguous [B,F,N] (graphML.py:4665-4667), one pass each way (magat_head_mean_forward / _backward)."""

    @staticmethod
    def forward(ctx, y_cat, P, F, relu):
        L = _cabi.lib()
        B, _, N = y_cat.shape
        yc = y_cat.detach().permute(0, 2, 1)
        if not yc.is_contiguous():
            yc = yc.contiguous()
        dev = y_cat.device
        with torch.cuda.device(dev):
            y = torch.empty((B, F, N), dtype=torch.float32, device=dev)
            _cabi.check(L.magat_head_mean_forward(yc.data_ptr(), B, N, P, F, int(relu), y.data_ptr(), _stream(dev)))
        ctx.dims = (B, N, P, F, bool(relu))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _cabi.lib()
        (y,) = ctx.saved_tensors
        B, N, P, F, relu = ctx.dims
        dev = y.device
        dy = dy.float().contiguous()
        with torch.cuda.device(dev):
            dyc = torch.empty((B, N, P * F), dtype=torch.float32, device=dev)
            _cabi.check(L.magat_head_mean_backward(dy.data_ptr(), y.data_ptr(), B, N, P, F, int(relu), dyc.data_ptr(),
                                                   _stream(dev)))
        return dyc.permute(0, 2, 1), None, None, None


def _apply_nonlinearity(fn, y, P, F, concatenate):
    """A custom nonlinearity sees what it sees in the reference: [B,P,F,N] before the heads are concatenated
    (graphML.py:4656-4662), [B,F,N] after the head mean (:4665-4667).  (ReLU is fused into the kernels.)"""
    if not concatenate:
        return fn(y)
    B, _, N = y.shape
    y4 = fn(y.permute(0, 2, 1).reshape(B, N, P, F).permute(0, 2, 3, 1))
    return y4.permute(0, 3, 1, 2).reshape(B, N, P * F).permute(0, 2, 1)


def _mode_of(attentionMode: str) -> int:
    if 'GAT_modified' in attentionMode:          # graphML.py:4588
        return _cabi.MODE_GAT_MODIFIED
    if attentionMode == 'KeyQuery':              # graphML.py:4594
        return _cabi.MODE_KEYQUERY
    raise ValueError(f"unsupported attentionMode {attentionMode!r}")


#: single-CTA-per-instance inference kernel: only for the simulator's shape, a handful of instances of <= 64
#: agents (measured: 0.08 ms at B=1 against 0.19 ms for the 9-launch path, but 2-7x slower than it from B=64 up,
#: where every CTA re-reads the same weight lines from L2)
_SMALL_N_MAX, _SMALL_B = 64, 4


class SparseAttention:
    """Attention of the last forward as the kernels keep it (sender-major lists); densified on demand."""

    def __init__(self, att, adj):
        self.att, self.adj = att, adj

    def dense(self):
        return attention_dense(self.att, self.adj)

    def dense_mean(self):
        return attention_dense(self.att, self.adj, mean_heads=True)


class DenseAttention:
    """Attention written densely by the small-graph kernel, [B,P,N,N]."""

    def __init__(self, aij):
        self.aij = aij

    def dense(self):
        return self.aij.unsqueeze(2)

    def dense_mean(self):
        return self.aij.mean(dim=1, keepdim=True)


def _small_forward(x, S, filterWeight, mixer, weight, weight_bias, bias, mode, concatenate, relu):
    """Inference-only single-launch path (magat_gat_forward_small)."""
    L = _cabi.lib()
    dev = x.device
    B, G, N = x.shape
    P, F, _, K, _ = filterWeight.shape
    S = S.detach()
    if S.dtype not in (torch.float32, torch.float64):
        S = S.to(torch.float32)
    if not S.is_contiguous():
        S = S.contiguous()
    xt = _node_major(x.detach())
    w, h = weight.detach().contiguous(), filterWeight.detach().contiguous()
    mx = None if mixer is None else mixer.detach().contiguous()
    wb = None if weight_bias is None else weight_bias.detach().contiguous()
    bs = None if bias is None else bias.detach().contiguous()
    with torch.cuda.device(dev):
        if concatenate:
            y_mem = torch.empty((B, N, P * F), dtype=torch.float32, device=dev)
            y = y_mem.permute(0, 2, 1)
        else:
            y_mem = torch.empty((B, F, N), dtype=torch.float32, device=dev)
            y = y_mem
        aij = torch.empty((B, P, N, N), dtype=torch.float32, device=dev)
        _cabi.check(L.magat_gat_forward_small(
            S.data_ptr(), _cabi.DT_F32 if S.dtype == torch.float32 else _cabi.DT_F64, xt.data_ptr(), _sb(xt), _sn(xt),
            w.data_ptr(), _p(mx), _p(wb), h.data_ptr(), _p(bs), y_mem.data_ptr(), y.stride(0), y.stride(2), y.stride(1),
            aij.data_ptr(), B, N, G, F, K, P, mode, int(concatenate), int(relu), _stream(dev)))
    return y, DenseAttention(aij)


def _fused_gso(x, S, G, F, K, P, mode, concatenate, path, max_degree, team=0):
    """The dense GSO wrapped for the fused forward, or None when that kernel does not cover the call."""
    B, N = x.shape[0], x.shape[2]
    if not (S.is_cuda and len(S.shape) == 4 and S.shape[0] == B and S.shape[1] == 1 and S.shape[2] == N
            and S.shape[3] == N):
        return None
    if path == "auto" and (not _FUSED_AUTO or N < _FUSED_MIN_N):
        return None
    D = min(32, (min(N, max_degree if max_degree else _FUSED_D0) + 3) // 4 * 4)
    if not _cabi.lib().magat_gat_fused_supported(N, G, F, K, P, D, mode, int(bool(concatenate))):
        return None
    if B * N * D * P >= 2 ** 31:
        return None
    S = S.detach()
    if S.dtype not in (torch.float32, torch.float64):
        S = S.to(torch.float32)
    if not S.is_contiguous():
        S = S.contiguous()
    xt = _node_major(x.detach())
    if S.data_ptr() % 16 or xt.data_ptr() % 16 or _sn(xt) % 4 or _sb(xt) % 4 or xt.dtype != torch.float32:
        return None
    return _FusedGSO(S, D, bool(max_degree), int(team or 0))


def gat_layer(x, S, filterWeight, mixer, weight, weight_bias, bias, *, mode: int, concatenate: bool,
              relu: bool = True, path: str = "auto", adjacency: Optional[Adjacency] = None,
              max_degree: Optional[int] = None, fused_team: int = 0):
    """One call of the fused layer.  Returns ``(y, attention)``; ``attention.dense()`` / ``.dense_mean()`` give
    ``aij`` [B,P,1,N,N] / its head mean on demand."""
    _require_cuda(x, "x")
    assert len(x.shape) == 3
    P, F, E, K, G = filterWeight.shape
    assert E == 1
    assert x.shape[1] == G                       # graphML.py:1735
    if x.dtype != torch.float32:
        x = x.float()
    _check_params(x, filterWeight, mixer if mode != _cabi.MODE_KEYQUERY else None, weight,
                  weight_bias if mode != _cabi.MODE_KEYQUERY else None, bias)
    B_, N_ = x.shape[0], x.shape[2]
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (x, filterWeight, mixer, weight, weight_bias, bias))
    if (adjacency is None and path == "auto" and not needs_grad and S is not None and S.is_cuda and S.shape[2] == N_
            and N_ <= _SMALL_N_MAX and B_ <= _SMALL_B
            and (mode != _cabi.MODE_KEYQUERY or F == G)
            and _cabi.lib().magat_gat_small_supported(N_, G, F, K, P, int(concatenate))):
        _require_cuda(S, "the GSO")
        assert len(S.shape) == 4 and S.shape[1] == 1 and S.shape[3] == N_ and S.shape[0] == B_
        return _small_forward(x, S, filterWeight, mixer if mode != _cabi.MODE_KEYQUERY else None, weight,
                              weight_bias if mode != _cabi.MODE_KEYQUERY else None, bias, mode, concatenate, relu)
    if (path == "auto" and (G < 128 or F < 128) and G <= 128 and F <= 128 and K <= 3
            and B_ * N_ >= _PAD_MIN_ROWS and (mode != _cabi.MODE_KEYQUERY or F == G)):
        return _padded_layer(x, S, filterWeight, mixer, weight, weight_bias, bias, mode, concatenate, relu, adjacency,
                             max_degree, fused_team)
    if (not concatenate and path in ("auto", "tcgen05", "fused") and F == 128 and G % 128 == 0 and K <= 3
            and B_ * N_ >= _MEAN_VIA_CONCAT_MIN_ROWS):
        # Heads AVERAGED (the reference's CLI default, main.py:113-115) at the tensor-core shapes: the per-head outputs
        # come from the concat path (tcgen05 projections forward and backward -- its weights do not fit TMEM for a
        # fused head sum); the mean over the heads, the ReLU and the reference's contiguous [B,F,N] layout
        # (graphML.py:4665-4667) follow in one transposing pass (_HeadMean), whose backward hands the layer its
        # per-head dY.  (B = 512, N = 1000, P = 4: 7.6 ms per training step against 60 ms on the generic fp32 kernels.)
        y_cat, att = gat_layer(x, S, filterWeight, mixer, weight, weight_bias, bias, mode=mode, concatenate=True,
                               relu=False, path=path, adjacency=adjacency, max_degree=max_degree, fused_team=fused_team)
        return _HeadMean.apply(y_cat, P, F, bool(relu)), att
    fused = None
    if adjacency is None and S is not None and not S.is_cuda:
        # the GSO stayed in host memory (where the reference's dataloader / simulator builds it): pack the mask there
        assert len(S.shape) == 4 and S.shape[0] == B_ and S.shape[2] == N_ and S.shape[3] == N_
        if S.shape[1] != 1:
            raise NotImplementedError("edge_features E != 1 is not supported")
        adjacency = build_adjacency_host(S, x.device, max_degree)
    if adjacency is None and path in ("auto", "fused") and S is not None:
        fused = _fused_gso(x, S, G, F, K, P, mode, concatenate, path, max_degree, fused_team)
    if path == "fused" and fused is None and adjacency is None:
        raise RuntimeError("path='fused': shape / layout not covered by the fused forward (magat_gat_fused_supported)")
    if fused is not None:
        adj = fused
    else:
        adj = adjacency if adjacency is not None else build_adjacency(S, max_degree)
        assert adj.B == x.shape[0] and adj.N == x.shape[2]
    if mode == _cabi.MODE_KEYQUERY:
        assert tuple(weight.shape) == (P, E, G, G)
        if F != G:
            raise RuntimeError("KeyQuery attention needs out_features == in_features "
                               "(the reference reshape at graphML.py:1765 fails otherwise)")
    else:
        assert tuple(weight.shape) == (P, E, F, G)
        assert tuple(mixer.shape) == (P, E, 2 * F)
    meta = _Meta()
    meta.mode, meta.concat, meta.relu, meta.path = mode, bool(concatenate), bool(relu), _PATH[path]
    meta.G, meta.F, meta.K, meta.P, meta.has_bias = G, F, K, P, bias is not None
    meta.needs_grad = needs_grad
    if mode == _cabi.MODE_KEYQUERY:
        mixer_in, wb_in = None, None             # unused by the math; grads stay None (graphML.py:1265-1266)
    else:
        mixer_in, wb_in = mixer, weight_bias
    y, att = _GATFunction.apply(x, weight, mixer_in, wb_in, filterWeight, bias, adj, meta)
    if fused is not None:
        adj = fused.adj
    return y, SparseAttention(att.detach(), adj)


def gat_layer_actions(x, S, filterWeight, mixer, weight, weight_bias, bias, head_weight, head_bias, *, mode: int,
                      relu: bool = True, adjacency: Optional[Adjacency] = None, max_degree: Optional[int] = None,
                      return_actions: bool = False):
    """SURVEY 8f row f3 (inference): the layer with concatenated heads AND the planner's linear action head
    (``actionsMLP`` = one ``nn.Linear(P*F, A)``, graphs/models/decentralplanner_GAT.py:221-233, :329-334) in one pass --
    ``y`` never reaches memory.  Returns ``logits [B*N, A]`` (row = b * N + n, as ``sharedFeature_stack`` orders them)
    and, with ``return_actions``, also ``argmax softmax`` per agent (utils/new_simulator.py:863-869) as int32 ``[B*N]``,
    plus the sparse attention of the call."""
    L = _cabi.lib()
    _require_cuda(x, "x")
    assert len(x.shape) == 3
    P, F, E, K, G = filterWeight.shape
    assert E == 1 and x.shape[1] == G
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in
                                       (x, filterWeight, mixer, weight, weight_bias, bias, head_weight, head_bias)):
        raise RuntimeError("gat_layer_actions is an inference call (nothing is kept for backward): use torch.no_grad()")
    if x.dtype != torch.float32:
        x = x.float()
    kq = mode == _cabi.MODE_KEYQUERY
    _check_params(x, filterWeight, None if kq else mixer, weight, None if kq else weight_bias, bias)
    A = head_weight.shape[0]
    assert tuple(head_weight.shape) == (A, P * F), "the head is nn.Linear(P*F, A) on the concatenated heads"
    for t, name in ((head_weight, "head_weight"), (head_bias, "head_bias")):
        if t is not None and (t.dtype != torch.float32 or t.device != x.device):
            raise RuntimeError(f"{name}: expected a float32 tensor on {x.device}")
    dev = x.device
    B, _, N = x.shape
    if adjacency is None and S is not None and not S.is_cuda:
        adjacency = build_adjacency_host(S, dev, max_degree)
    adj = adjacency if adjacency is not None else build_adjacency(S, max_degree)
    assert adj.B == B and adj.N == N
    D = adj.D
    xt = _node_major(x.detach())
    weight_c, filt_c = weight.detach().contiguous(), filterWeight.detach().contiguous()
    mixer_c = None if kq or mixer is None else mixer.detach().contiguous()
    wb_c = None if kq or weight_bias is None else weight_bias.detach().contiguous()
    bias_c = None if bias is None else bias.detach().contiguous()
    hw_c = head_weight.detach().contiguous()
    hb_c = None if head_bias is None else head_bias.detach().contiguous()
    with torch.cuda.device(dev):
        att = torch.empty((B, N, D, P), dtype=torch.float32, device=dev)
        taps = torch.empty((B, N, P, max(K - 1, 1), G), dtype=torch.float32, device=dev) if K > 1 else None
        wprep = torch.empty(L.magat_gat_wprep_floats(G, F, K, P, mode), dtype=torch.float32, device=dev)
        sproj = torch.empty((B, N, P, G if kq else 2), dtype=torch.float32, device=dev)
        partial = torch.empty((P, B * N, 8), dtype=torch.float32, device=dev)
        logits = torch.empty((B * N, A), dtype=torch.float32, device=dev)
        actions = torch.empty((B * N,), dtype=torch.int32, device=dev) if return_actions else None
        a = _cabi.FwdArgs(B=B, N=N, G=G, F=F, K=K, P=P, D=D, mode=mode, concat=1, relu=int(relu),
                          path=_cabi.PATH_AUTO, reserved=0, x=xt.data_ptr(), x_sb=_sb(xt), x_sn=_sn(xt),
                          nbr_out=adj.nbr_out.data_ptr(), nbr_in=adj.nbr_in.data_ptr(), slot_in=adj.slot_in.data_ptr(),
                          slot_out=_p(adj.slot_out), weight=weight_c.data_ptr(), mixer=_p(mixer_c), weight_bias=_p(wb_c),
                          filterWeight=filt_c.data_ptr(), bias=_p(bias_c), y=None, y_sb=N * P * F, y_sn=P * F, y_sc=1,
                          att=att.data_ptr(), ain=None, taps=_p(taps), wprep=wprep.data_ptr(), sproj=sproj.data_ptr(),
                          relu_bits=None)
        if not L.magat_gat_actions_supported(a, A):
            raise RuntimeError("gat_layer_actions: needs F = 128, G a multiple of 128, K <= 3 and at most 8 actions "
                               "(magat_gat_actions_supported); use the layer followed by the nn.Linear otherwise")
        _cabi.check(L.magat_gat_forward_actions(a, hw_c.data_ptr(), _p(hb_c), A, partial.data_ptr(), logits.data_ptr(),
                                                _p(actions), _stream(dev)))
    att_out = SparseAttention(att, adj)
    return (logits, actions, att_out) if return_actions else (logits, att_out)


#: heads averaged: from this many node rows on the per-head outputs come from the concat path (below, the two extra
#: launches of the head mean cost more than the generic backward kernels do)
_MEAN_VIA_CONCAT_MIN_ROWS = 2048

#: from this many node rows (B * N) on, layers with fewer than 128 features are zero padded onto the 128-feature kernels
_PAD_MIN_ROWS = 32768


def _padded_layer(x, S, filterWeight, mixer, weight, weight_bias, bias, mode, concatenate, relu, adjacency, max_degree,
                  fused_team):
    """Fewer than 128 in / out features (the published bottleneck model has 32, README.md:385-396 of the reference) at
    scale: the tcgen05 projections and the lean sparse kernels are built around 128 features, the generic fp32 kernels
    that take any width are several times slower per byte.  Zero padding x, the parameters and (implicitly) y to 128
    features changes nothing in the math -- padded inputs meet zero weights, padded outputs are relu(0) and are cut off
    -- and runs 2-4x faster despite the larger tensors (B = 128, N = 1000: G = F = 32 2.9 -> 1.3 ms per training step,
    G = F = 64 7.8 -> 1.9 ms).  Differentiable torch pads / slices: autograd cuts the gradients back."""
    pad = torch.nn.functional.pad
    P, F, E, K, G = filterWeight.shape
    B, N = x.shape[0], x.shape[2]
    pg, pf = 128 - G, 128 - F
    if mode == _cabi.MODE_KEYQUERY:
        assert tuple(weight.shape) == (P, E, G, G)
    else:
        assert tuple(weight.shape) == (P, E, F, G) and tuple(mixer.shape) == (P, E, 2 * F)
    xp = pad(x.permute(0, 2, 1), (0, pg)).permute(0, 2, 1)               # [B,128,N] over node-major memory
    fw = pad(filterWeight, (0, pg, 0, 0, 0, 0, 0, pf))                   # [P,128,E,K,128]
    if mode == _cabi.MODE_KEYQUERY:
        w, mx, wb = pad(weight, (0, pg, 0, pg)), mixer, weight_bias      # [P,E,128,128]; mixer / weight_bias unused
    else:
        w = pad(weight, (0, pg, 0, pf))                                  # [P,E,128,128]
        mx = torch.cat((pad(mixer[..., :F], (0, pf)), pad(mixer[..., F:], (0, pf))), dim=-1)     # a1 | a2
        wb = pad(weight_bias, (0, pf))
    bp = None if bias is None else pad(bias, (0, 0, 0, pf))
    y, att = gat_layer(xp, S, fw, mx, w, wb, bp, mode=mode, concatenate=concatenate, relu=relu, path="auto",
                       adjacency=adjacency, max_degree=max_degree, fused_team=fused_team)
    if concatenate:
        y = y.permute(0, 2, 1).reshape(B, N, P, 128)[..., :F].reshape(B, N, P * F).permute(0, 2, 1)
    else:
        y = y[:, :F, :].contiguous()
    return y, att


def attention_dense(att: torch.Tensor, adj: Adjacency, mean_heads: bool = False) -> torch.Tensor:
    """Sparse attention -> dense ``aij`` [B,P,1,N,N] (or its head mean [B,1,N,N])."""
    L = _cabi.lib()
    B, N, D, P = att.shape
    dev = att.device
    with torch.cuda.device(dev):
        out = torch.zeros((B, 1, N, N) if mean_heads else (B, P, 1, N, N), dtype=torch.float32, device=dev)
        _cabi.check(L.magat_gat_attention_dense(att.data_ptr(), adj.nbr_out.data_ptr(), B, N, D, P,
                                                int(mean_heads), out.data_ptr(), _stream(dev)))
    return out


# ---- functionals with the reference's signatures -------------------------------------------------

def _functional(h, x, a, W, W_b, S, b, mode):
    P, F = h.shape[0], h.shape[1]
    B, N = x.shape[0], x.shape[2]
    y, att = gat_layer(x, S, h, a, W, W_b, b, mode=mode, concatenate=True, relu=False)
    y = y.permute(0, 2, 1).reshape(B, N, P, F).permute(0, 2, 3, 1)      # B x P x F x N
    return y, att.dense()


def _check_slope(negative_slope):
    if negative_slope != 0.2:
        raise NotImplementedError("only the reference's default negative_slope = 0.2 is built in "
                                  "(graphML.py:713; no caller overrides it)")


def graphAttentionLSIGFBatch_KeyQuery(h, x, a, W, W_b, S, b=None, negative_slope=0.2):
    """graphML.py:1724-1775: returns (y [B,P,F,N] before the nonlinearity, aij [B,P,E,N,N])."""
    _check_slope(negative_slope)
    return _functional(h, x, a, W, W_b, S, b, _cabi.MODE_KEYQUERY)


def graphAttentionLSIGFBatch_modified(h, x, a, W, W_b, S, b=None, negative_slope=0.2):
    """graphML.py:1777-1827."""
    _check_slope(negative_slope)
    return _functional(h, x, a, W, W_b, S, b, _cabi.MODE_GAT_MODIFIED)


def _attention_only(x, a, W, W_b, S, mode):
    P, E, G = W.shape[0], W.shape[1], W.shape[3]
    F = G if mode == _cabi.MODE_KEYQUERY else W.shape[2]
    h = torch.zeros((P, F, E, 1, G), dtype=torch.float32, device=x.device)
    with torch.no_grad():
        _, att = gat_layer(x, S, h, a, W, W_b, None, mode=mode, concatenate=True, relu=False)
    return att.dense()


def learnAttentionGSOBatch_KeyQuery(x, a, W, W_b, S, negative_slope=0.2):
    """graphML.py:1180-1286, same argument order (the reference ignores ``a``, ``W_b`` and ``negative_slope`` in
    this mode: no LeakyReLU on the scores, :1265-1266)."""
    return _attention_only(x, a, W, None, S, _cabi.MODE_KEYQUERY)


def learnAttentionGSOBatch(x, a, W, W_b, S, negative_slope=0.2):
    """graphML.py:713-823."""
    _check_slope(negative_slope)
    return _attention_only(x, a, W, W_b, S, _cabi.MODE_GAT_MODIFIED)


# ---- SURVEY 8f row f4 (first of the ablation modes): GAT_origin on the same kernels -----------------------------
# learnAttentionGSOBatch_origin (graphML.py:964-1070) is the GAT_modified score WITHOUT the per-head bias on
# z = W x and with a self loop added to the GSO (S + I, :1019) before the edge test; graphAttentionLSIGFBatch_Origin
# (:1939-2005) filters with K SCALAR taps h[e, k] times a reshaped copy of the attention weight W (:1964-1969).  Both
# are expressed here on the GAT_modified kernels: the effective per-head filter h[k] * W' is a differentiable torch
# expression in front of the fused layer (autograd carries dh and the filter part of dW), the self loops are one
# elementwise pass over S.

def _origin_gso(S: torch.Tensor) -> torch.Tensor:
    B, E, N, _ = S.shape
    eye = torch.eye(N, dtype=torch.float32, device=S.device)
    return S.to(torch.float32) + eye            # fp32 as the reference casts it (:1019); a -1 diagonal cancels the loop


def _origin_filter(h: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    P, E, F, G = W.shape
    K = h.shape[1]
    Wr = W.permute(0, 3, 1, 2).reshape(P, F, E, 1, G)           # the reference's own reshape (:1966-1967), kept as is
    return h.reshape(1, 1, E, K, 1) * Wr                           # P x F x E x K x G


def _origin_layer(x, S, h, a, W, b, concatenate, relu, path="auto", max_degree=None):
    P, E, F, G = W.shape
    if E != 1:
        raise NotImplementedError("edge_features E != 1 is not supported")
    wb0 = torch.zeros((P, E, F), dtype=torch.float32, device=W.device)
    _require_cuda(S, "the GSO")
    if S.shape[2] <= _SMALL_N_MAX and S.shape[0] <= _SMALL_B:
        # the simulator's shape: S + I is a handful of kilobytes and the single-launch small-graph kernel takes it dense
        return gat_layer(x, _origin_gso(S), _origin_filter(h, W), a, W, wb0, b, mode=_cabi.MODE_GAT_MODIFIED,
                         concatenate=concatenate, relu=relu, path=path, max_degree=max_degree)
    adj = build_adjacency(S, max_degree, self_loops=True)       # the edges of S + I without materialising it
    return gat_layer(x, None, _origin_filter(h, W), a, W, wb0, b, mode=_cabi.MODE_GAT_MODIFIED,
                     concatenate=concatenate, relu=relu, path="auto" if path == "fused" else path, adjacency=adj,
                     max_degree=max_degree)


def learnAttentionGSOBatch_origin(x, a, W, S, negative_slope=0.2):
    """graphML.py:964-1070."""
    _check_slope(negative_slope)
    P, E, F, G = W.shape
    wb0 = torch.zeros((P, E, F), dtype=torch.float32, device=W.device)
    return _attention_only(x, a, W, wb0, _origin_gso(S), _cabi.MODE_GAT_MODIFIED)     # (functional: S + I as a tensor)


def graphAttentionLSIGFBatch_Origin(h, x, a, W, S, b=None, negative_slope=0.2):
    """graphML.py:1939-2005: returns (y [B,P,F,N] before the nonlinearity, aij [B,P,E,N,N])."""
    _check_slope(negative_slope)
    P, E, F, G = W.shape
    B, N = x.shape[0], x.shape[2]
    y, att = _origin_layer(x, S, h, a, W, b, True, False)
    return y.permute(0, 2, 1).reshape(B, N, P, F).permute(0, 2, 3, 1), att.dense()


class GraphFilterBatchAttentional_Origin(nn.Module):
    """Mirror of the reference module of the same name (graphML.py:4175-4339): the GAT_origin ablation
    (``--attentionMode GAT_origin``, graphs/models/decentralplanner_GAT.py:186-189).  Parameters, their order and
    initialisation as in the reference: mixer [P,E,2F], weight [P,E,F,G], filterWeight [E,K], bias [F,1]."""

    def __init__(self, G, F, K, P, E=1, bias=True, nonlinearity=nn.functional.relu, concatenate=True,
                 attentionMode='GAT_origin'):
        super().__init__()
        self.G, self.F, self.K, self.P, self.E = G, F, K, P, E
        self.S = None
        self._last, self._aij = None, None
        self.nonlinearity = nonlinearity
        self.concatenate = concatenate
        self.attentionMode = attentionMode
        self.path = "auto"
        self.mixer = nn.parameter.Parameter(torch.Tensor(P, E, 2 * F))
        self.weight = nn.parameter.Parameter(torch.Tensor(P, E, F, G))
        self.filterWeight = nn.parameter.Parameter(torch.Tensor(E, K))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.P)
        for p_ in (self.weight, self.mixer, self.filterWeight):
            p_.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        assert len(S.shape) == 4
        assert S.shape[1] == self.E
        self.N = S.shape[2]
        assert S.shape[3] == self.N
        self.S = S

    @property
    def aij(self):
        if self._aij is None and self._last is not None:
            self._aij = self._last.dense().cpu().numpy()
        return self._aij

    @aij.setter
    def aij(self, value):
        self._aij = value

    def returnAttentionGSO(self):
        aij = self.aij
        assert len(aij.shape) == 5
        assert aij.shape[2] == self.E
        self.N = aij.shape[3]
        return np.mean(aij, axis=1)

    def forward(self, x):
        B, F, Nin = x.shape
        if Nin < self.N:
            x = torch.cat((x, torch.zeros(B, F, self.N - Nin).type(x.dtype).to(x.device)), dim=2)
        if self.S is None:
            raise RuntimeError("GraphFilterBatchAttentional_Origin.forward: no GSO stored -- call addGSO(S) first")
        fused_relu = self.nonlinearity in (nn.functional.relu, torch.relu)
        y, att = _origin_layer(x, self.S, self.filterWeight, self.mixer, self.weight, self.bias, self.concatenate,
                               fused_relu, path=self.path, max_degree=getattr(self, "max_degree", None))
        self._last, self._aij = att, None
        if not fused_relu:
            y = _apply_nonlinearity(self.nonlinearity, y, self.P, self.F, self.concatenate)
        if Nin < self.N:
            y = torch.index_select(y, 2, torch.arange(Nin).to(y.device))
        return y

    def extra_repr(self):
        rep = "in_features=%d, out_features=%d, filter_taps=%d, attention_heads=%d, edge_features=%d, bias=%s, " % (
            self.G, self.F, self.K, self.P, self.E, self.bias is not None)
        rep += "attentionMode=%s, " % (self.attentionMode)
        rep += ("GSO stored: number_nodes=%d" % self.N) if self.S is not None else "no GSO stored"
        return rep

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_last"] = None
        state["_aij"] = None
        return state


# ---- the non-attentional graph filter on the same kernels ------------------------------------

class _LSIGFFunction(torch.autograd.Function):
    """BatchLSIGF (graphML.py:5485-5579): y = sum_k H_k (x S^k) + b.  The tap recursion and the K-tap projection are
    the attention layer's kernels with one head whose "attention" is the GSO's own edge values."""

    @staticmethod
    def forward(ctx, x, weight, bias, adj: Adjacency, vals, path):
        L = _cabi.lib()
        dev = x.device
        B, G, N = x.shape
        F, _, K, _ = weight.shape
        D = adj.D
        xt = _node_major(x.detach())
        filt_c = weight.detach().contiguous()                   # [F,1,K,G] is [P=1][F][1][K][G]
        bias_c = None if bias is None else bias.detach().contiguous()
        with torch.cuda.device(dev):
            y_mem = torch.empty((B, N, F), dtype=torch.float32, device=dev)
            y = y_mem.permute(0, 2, 1)                          # the reference returns this permuted view too (:5573-5575)
            ain = None
            taps = torch.empty((B, N, 1, max(K - 1, 1), G), dtype=torch.float32, device=dev) if K > 1 else None
            wprep = torch.empty(L.magat_gat_wprep_floats(G, F, K, 1, _cabi.MODE_GSO_VALUES), dtype=torch.float32, device=dev)
            dummy = torch.empty(4, dtype=torch.float32, device=dev)
            a = _cabi.FwdArgs(B=B, N=N, G=G, F=F, K=K, P=1, D=D, mode=_cabi.MODE_GSO_VALUES, concat=1, relu=0,
                              path=_PATH[path], reserved=0, x=xt.data_ptr(), x_sb=_sb(xt), x_sn=_sn(xt),
                              nbr_out=adj.nbr_out.data_ptr(), nbr_in=adj.nbr_in.data_ptr(),
                              slot_in=adj.slot_in.data_ptr(), slot_out=_p(adj.slot_out),
                              weight=filt_c.data_ptr(), mixer=None, weight_bias=None,
                              filterWeight=filt_c.data_ptr(), bias=_p(bias_c),
                              y=y_mem.data_ptr(), y_sb=y.stride(0), y_sn=y.stride(2), y_sc=y.stride(1),
                              att=vals.data_ptr(), ain=_p(ain), taps=_p(taps), wprep=wprep.data_ptr(),
                              sproj=dummy.data_ptr())
            _cabi.check(L.magat_gat_forward(a, _stream(dev)))
            ctx.taps_valid = L.magat_gat_forward_taps_valid(a)
        ctx.adj, ctx.path, ctx.has_bias = adj, path, bias is not None
        ctx.save_for_backward(xt, filt_c, y, vals, taps, wprep)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _cabi.lib()
        adj = ctx.adj
        xt, filt_c, y, vals, taps, wprep = ctx.saved_tensors
        dev = xt.device
        B, N, G = xt.shape
        F, _, K, _ = filt_c.shape
        D = adj.D
        need_dx, need_df, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and ctx.has_bias
        if dy.dtype != torch.float32:
            dy = dy.float()
        with torch.cuda.device(dev):
            def buf(*shape):
                return torch.empty(shape, dtype=torch.float32, device=dev)
            dx = buf(B, N, G) if need_dx else None
            df = torch.empty_like(filt_c) if need_df else None
            db = buf(F, 1) if need_db else None
            dummy = buf(4)
            nws = L.magat_gat_backward_workspace_bytes(B, N, G, F, K, 1, D, _cabi.MODE_GSO_VALUES)
            ws = torch.empty(nws + 256, dtype=torch.uint8, device=dev)
            ws = ws[(-ws.data_ptr()) % 256:]
            a = _cabi.BwdArgs(B=B, N=N, G=G, F=F, K=K, P=1, D=D, mode=_cabi.MODE_GSO_VALUES, concat=1, relu=0,
                              path=_PATH[ctx.path], need_dx=int(need_dx), need_dweight=0, need_dfilter=int(need_df),
                              need_dbias=int(need_db), need_dmixer=0, taps_valid=int(ctx.taps_valid), reserved=0,
                              x=xt.data_ptr(), x_sb=_sb(xt), x_sn=_sn(xt),
                              nbr_out=adj.nbr_out.data_ptr(), nbr_in=adj.nbr_in.data_ptr(), slot_in=adj.slot_in.data_ptr(),
                              weight=filt_c.data_ptr(), mixer=None, weight_bias=None, filterWeight=filt_c.data_ptr(),
                              y=y.data_ptr(), y_sb=y.stride(0), y_sn=y.stride(2), y_sc=y.stride(1),
                              att=vals.data_ptr(), taps=_p(taps), wprep=wprep.data_ptr(), sproj=dummy.data_ptr(),
                              dy=dy.data_ptr(), dy_sb=dy.stride(0), dy_sn=dy.stride(2), dy_sc=dy.stride(1),
                              dx=_p(dx), dweight=None, dmixer=None, dweight_bias=None,
                              dfilterWeight=_p(df), dbias=_p(db),
                              gz=None, datt=None, rc=None, partial=None)
            _cabi.check(L.magat_gat_backward_ws(a, ws.data_ptr(), nws, _stream(dev)))
        return (dx.permute(0, 2, 1) if need_dx else None), df, db, None, None, None


def BatchLSIGF(h, S, x, b=None):
    """graphML.py:5485-5579, same signature: h [F,E,K,G], S [B,E,N,N], x [B,G,N], b [F,1] -> y [B,F,N]."""
    return _lsigf(h, S, x, b)


def _lsigf(h, S, x, b=None, path="auto", max_degree=None):
    _require_cuda(x, "x")
    F, E, K, G = h.shape
    assert S.shape[1] == E
    N = S.shape[2]
    assert S.shape[3] == N
    B = x.shape[0]
    assert x.shape[1] == G
    assert x.shape[2] == N
    if E != 1:
        raise NotImplementedError("edge_features E != 1 is not supported (the planners use E = 1)")
    if b is not None and (b.shape[-1] != 1):
        raise NotImplementedError("per-node bias [F,N] is not supported (the planners use [F,1])")
    if x.dtype != torch.float32:
        x = x.float()
    _check_params(x, h, None, None, None, b)
    adj = build_adjacency(S, max_degree, nonzero=True)
    Sd = S.detach()
    if Sd.dtype not in (torch.float32, torch.float64):
        Sd = Sd.to(torch.float32)
    if not Sd.is_contiguous():
        Sd = Sd.contiguous()
    dev = x.device
    with torch.cuda.device(dev):
        vals = torch.empty((B, N, adj.D, 1), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib().magat_gso_edge_values(
            Sd.data_ptr(), _cabi.DT_F32 if Sd.dtype == torch.float32 else _cabi.DT_F64, adj.nbr_out.data_ptr(), B, N, adj.D,
            vals.data_ptr(), _stream(dev)))
    return _LSIGFFunction.apply(x, h, b, adj, vals, path)


class GraphFilterBatch(nn.Module):
    """Drop-in for ``utils.graphUtils.graphML.GraphFilterBatch`` (graphML.py:5581-5700), the graph convolution of the
    GNN baseline planners (graphs/models/decentralplanner.py:280): same constructor, parameters (``weight`` [F,E,K,G],
    ``bias`` [F,1]), ``addGSO``, ``forward`` and ``extra_repr``.  No nonlinearity inside (the planner appends its own)."""

    def __init__(self, G, F, K, E=1, bias=True):
        super().__init__()
        self.G = G
        self.F = F
        self.K = K
        self.E = E
        self.S = None
        self.path = "auto"
        self.max_degree = None
        if E != 1:
            raise NotImplementedError("edge_features E != 1 is not supported")
        self.weight = nn.parameter.Parameter(torch.Tensor(F, E, K, G))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.K)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        assert len(S.shape) == 4
        assert S.shape[1] == self.E
        self.N = S.shape[2]
        assert S.shape[3] == self.N
        self.S = S

    def forward(self, x):
        B = x.shape[0]
        F = x.shape[1]
        Nin = x.shape[2]
        if Nin < self.N:
            x = torch.cat((x, torch.zeros(B, F, self.N - Nin).type(x.dtype).to(x.device)), dim=2)
        u = _lsigf(self.weight, self.S, x, self.bias, path=self.path, max_degree=self.max_degree)
        if Nin < self.N:
            u = torch.index_select(u, 2, torch.arange(Nin).to(u.device))
        return u

    def extra_repr(self):
        reprString = "in_features=%d, out_features=%d, " % (self.G, self.F) + "filter_taps=%d, " % (self.K) + \
            "edge_features=%d, " % (self.E) + "bias=%s, " % (self.bias is not None)
        if self.S is not None:
            reprString += "GSO stored"
        else:
            reprString += "no GSO stored"
        return reprString


# ---- the module ------------------------------------------------------------------------------

class GraphFilterBatchAttentional(nn.Module):
    """Drop-in for ``utils.graphUtils.graphML.GraphFilterBatchAttentional`` (graphML.py:4506-4685).

    Same constructor, parameter names / shapes / registration order (so reference checkpoints
    ``GFL.{l}.*`` load), ``addGSO``, ``forward``, ``returnAttentionGSO`` and ``extra_repr``.
    Differences, all invisible to the planners: the work runs in sm_100a CUDA kernels, and
    ``self.aij`` (a dense numpy copy the reference makes on every forward, graphML.py:4650) is
    materialised lazily on first access.
    """

    def __init__(self, G, F, K, P, E=1, bias=True, nonlinearity=nn.functional.relu, concatenate=True,
                 attentionMode='GAT_modified'):
        super().__init__()
        self.G = G
        self.F = F
        self.K = K
        self.P = P
        self.E = E
        self.S = None
        self.nonlinearity = nonlinearity
        self.concatenate = concatenate
        self.attentionMode = attentionMode
        self.path = "auto"
        #: optional promise that no agent has more than this many in- or out-neighbours: lets the fused forward size
        #: its neighbour lists without reading the degree statistics back (no host synchronisation in forward)
        self.max_degree = None
        #: CTAs per planning instance of the fused forward (path="fused"): 0 = default (8), or 16
        self.fused_team = 0
        self._last = None
        self._aij = None
        self._adj = None
        if E != 1:
            raise NotImplementedError("edge_features E != 1 is not supported")
        self.mixer = nn.parameter.Parameter(torch.Tensor(P, E, 2 * F))
        self.weight_bias = nn.parameter.Parameter(torch.Tensor(P, E, F))
        self.filterWeight = nn.parameter.Parameter(torch.Tensor(P, F, E, K, G))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter('bias', None)
        if 'GAT_modified' in attentionMode:
            self.weight = nn.parameter.Parameter(torch.Tensor(P, E, F, G))
        elif attentionMode == 'KeyQuery':
            self.weight = nn.parameter.Parameter(torch.Tensor(P, E, G, G))
        self.reset_parameters()          # raises AttributeError for other modes, like the reference

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.P)
        self.weight.data.uniform_(-stdv, stdv)
        self.weight_bias.data.uniform_(0, 0)
        self.mixer.data.uniform_(-stdv, stdv)
        self.filterWeight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        assert len(S.shape) == 4
        assert S.shape[1] == self.E
        self.N = S.shape[2]
        assert S.shape[3] == self.N
        self.S = S                       # borrowed, read at forward time like the reference
        self._adj = None

    def addAdjacency(self, adj: Adjacency):
        """Not in the reference: hand the layer neighbour lists built ahead of time (``build_adjacency``,
        ``build_adjacency_host``, ``build_adjacency_from_rowbits`` -- e.g. by a prefetching thread on a side stream)."""
        self.N = adj.N
        self.S = None
        self._adj = adj

    def addGSOFromPositions(self, pos, comm_radius):
        """Not in the reference: give the layer the agents' positions [B,N,2] instead of
        the dense GSO the simulator derives from them (utils/new_simulator.py:823-827).  The neighbour lists are built
        on the device, no N x N matrix exists anywhere, and 8N bytes per instance cross PCIe instead of 4N^2."""
        assert len(pos.shape) == 3 and pos.shape[2] == 2
        self.N = pos.shape[1]
        self.S = None
        self._adj = build_adjacency_from_positions(pos, comm_radius, getattr(self, "max_degree", None))

    # ``aij`` is what graphML.py:4650 stores eagerly; here the dense copy is made on first use.
    @property
    def aij(self):
        if self._aij is None and self._last is not None:
            self._aij = self._last.dense().cpu().numpy()
        return self._aij

    @aij.setter
    def aij(self, value):
        self._aij = value
        self._last = None

    def returnAttentionGSO(self):
        if self._aij is None and self._last is not None:
            return self._last.dense_mean().cpu().numpy()
        aij = self.aij
        assert len(aij.shape) == 5
        assert aij.shape[2] == self.E
        self.N = aij.shape[3]
        return np.mean(aij, axis=1)

    def forward(self, x):
        B = x.shape[0]
        F = x.shape[1]
        Nin = x.shape[2]
        if Nin < self.N:                 # graphML.py:4642-4646
            x = torch.cat((x, torch.zeros(B, F, self.N - Nin).type(x.dtype).to(x.device)), dim=2)
        fused_relu = self.nonlinearity in (nn.functional.relu, torch.relu)
        if self.S is None and getattr(self, "_adj", None) is None:
            raise RuntimeError("GraphFilterBatchAttentional.forward: no GSO stored -- call addGSO(S) (or "
                               "addGSOFromPositions) first; device-side neighbour lists do not survive pickling")
        y, att = gat_layer(x, self.S, self.filterWeight, self.mixer, self.weight, self.weight_bias,
                           self.bias, mode=_mode_of(self.attentionMode), concatenate=self.concatenate,
                           relu=fused_relu, path=self.path, adjacency=getattr(self, "_adj", None),
                           max_degree=getattr(self, "max_degree", None), fused_team=getattr(self, "fused_team", 0))
        self._last, self._aij = att, None
        if not fused_relu:
            y = _apply_nonlinearity(self.nonlinearity, y, self.P, self.F, self.concatenate)
        if Nin < self.N:
            y = torch.index_select(y, 2, torch.arange(Nin).to(y.device))
        return y

    def forward_actions(self, x, actions_mlp, return_actions: bool = False):
        """``actions_mlp(self(x).permute(0, 2, 1).reshape(B * N, -1))`` in one pass, for inference
        (``torch.no_grad()``), when ``actions_mlp`` is the reference's single ``nn.Linear`` (optionally wrapped in an
        ``nn.Sequential``; graphs/models/decentralplanner_GAT.py:221-237) and the heads are concatenated: the layer's
        output is never written.  Anything else (more layers, dropout in training mode, head mean, a custom
        nonlinearity, padded inputs) takes the two-step route, with the same result."""
        lin = actions_mlp
        if isinstance(lin, nn.Sequential):
            mods = [m for m in lin if not (isinstance(m, nn.Dropout) and not m.training)]
            lin = mods[0] if len(mods) == 1 else None
        fusable = (isinstance(lin, nn.Linear) and self.concatenate and x.shape[2] == self.N
                   and self.nonlinearity in (nn.functional.relu, torch.relu) and not torch.is_grad_enabled()
                   and lin.in_features == self.P * self.F and lin.out_features <= 8 and self.F == 128
                   and self.G % 128 == 0 and self.K <= 3 and self.path in ("auto", "tcgen05"))
        if self.S is None and getattr(self, "_adj", None) is None:
            raise RuntimeError("GraphFilterBatchAttentional.forward_actions: no GSO stored -- call addGSO(S) first")
        if not fusable:
            y = self.forward(x)
            logits = actions_mlp(y.permute(0, 2, 1).reshape(y.shape[0] * y.shape[2], -1))
            return (logits, torch.max(logits, 1)[1].int()) if return_actions else logits
        out = gat_layer_actions(x, self.S, self.filterWeight, self.mixer, self.weight, self.weight_bias, self.bias,
                                lin.weight, lin.bias, mode=_mode_of(self.attentionMode), relu=True,
                                adjacency=getattr(self, "_adj", None), max_degree=getattr(self, "max_degree", None),
                                return_actions=return_actions)
        self._last, self._aij = out[-1], None
        return (out[0], out[1]) if return_actions else out[0]

    def extra_repr(self):
        reprString = "in_features=%d, " % self.G
        reprString += "out_features=%d, " % self.F
        reprString += "filter_taps=%d, " % self.K
        reprString += "attention_heads=%d, " % self.P
        reprString += "edge_features=%d, " % self.E
        reprString += "bias=%s, " % (self.bias is not None)
        reprString += "attentionMode=%s, " % (self.attentionMode)
        if self.S is not None:
            reprString += "GSO stored: number_nodes=%d" % (self.N)
        else:
            reprString += "no GSO stored"
        return reprString

    def __getstate__(self):
        # picklable for torch.multiprocessing.spawn (agents/...GAT.py:720-728): no device scratch, no handles
        state = self.__dict__.copy()
        state["_last"] = None
        state["_aij"] = None
        state["_adj"] = None
        return state
