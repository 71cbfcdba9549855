"""The C-ABI library loads and exports every symbol include/magat_gat.h declares (no GPU needed),
and the host-side mirror keeps the reference's module surface."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "magat_gat.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from magat_pathplanning_b200 import _cabi
    return _cabi


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(magat_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    names = declared_functions()
    assert len(names) >= 10
    L = ctypes.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/magat_gat.h but not exported"
    assert set(names) == set(built.EXPORTS)
    assert L.magat_abi_version() == built.ABI_VERSION


def test_struct_sizes_match_header(built):
    # 12 x int32 + 21 x 8 bytes ; 18 x int32 + 32 x 8 bytes
    assert ctypes.sizeof(built.FwdArgs) == 12 * 4 + 22 * 8
    assert ctypes.sizeof(built.BwdArgs) == 18 * 4 + 33 * 8
    # 14 x int32 + 23 x 8 bytes
    assert ctypes.sizeof(built.FusedArgs) == 14 * 4 + 23 * 8


def test_sizes_helpers_need_no_gpu(built):
    L = built.lib()
    assert L.magat_gat_wprep_floats(128, 128, 3, 4, built.MODE_GAT_MODIFIED) >= 4 * 2 * 128 + 8
    assert L.magat_gat_bwd_partial_floats(512, 1000, 128, 128, 3, 4, built.MODE_KEYQUERY) > 0


def test_bad_arguments_return_codes_not_crash(built):
    L = built.lib()
    a = built.FwdArgs(B=1, N=4, G=8, F=16, K=2, P=1, D=1, mode=built.MODE_KEYQUERY)
    rc = L.magat_gat_forward(a, None)
    assert rc == 2 and b"F == G" in L.magat_last_error()          # MAGAT_E_UNSUPPORTED
    a = built.FwdArgs(B=0, N=4, G=8, F=8, K=2, P=1, D=1, mode=built.MODE_KEYQUERY)
    assert L.magat_gat_forward(a, None) == 1                        # MAGAT_E_BAD_ARG
    assert L.magat_gat_forward(None, None) == 1
    assert L.magat_gso_scan(None, 0, 1, 4, None, None, None, None) == 1


def test_fused_entry_point_rejects_without_crashing(built):
    L = built.lib()
    assert L.magat_gat_forward_fused(None, None) == 1
    a = built.FusedArgs(B=2, N=100, G=128, F=128, K=3, P=4, D=16, mode=built.MODE_KEYQUERY, concat=0)
    assert L.magat_gat_forward_fused(a, None) == 2 and b"not covered" in L.magat_last_error()
    assert L.magat_gat_fused_supported(1000, 128, 128, 3, 4, 16, built.MODE_KEYQUERY, 1) == 1
    assert L.magat_gat_fused_supported(1000, 128, 128, 3, 4, 36, built.MODE_KEYQUERY, 1) == 0     # D > 32
    assert L.magat_gat_fused_supported(1001, 128, 128, 3, 4, 16, built.MODE_KEYQUERY, 1) == 0     # N % 4
    assert L.magat_gat_fused_supported(1000, 64, 64, 3, 4, 16, built.MODE_KEYQUERY, 1) == 0
    n = L.magat_gat_fused_workspace_bytes(512, 1000, 3, 4, 16, built.MODE_KEYQUERY, 0, 0)
    assert 0 < n < 1 << 30
    assert L.magat_gat_fused_workspace_bytes(512, 1000, 3, 4, 16, built.MODE_KEYQUERY, 1, 0) < n
    assert L.magat_gat_fused_workspace_bytes(512, 1000, 3, 4, 16, built.MODE_KEYQUERY, 0, 16) < n       # fewer teams
    assert L.magat_gat_fused_workspace_bytes(512, 1000, 3, 4, 16, built.MODE_KEYQUERY, 0, 12) == 0


def test_functional_signatures_match_reference():
    """The functionals are rebound by name inside the reference module (integration.install_into_reference): same
    parameter names, order and defaults (graphML.py:713, :1180, :1724, :1777)."""
    import inspect
    from oracle.ref_loader import load_reference_graphml, reference_available
    if not reference_available():
        pytest.skip("reference checkout not present")
    gml = load_reference_graphml()
    import magat_pathplanning_b200 as ours
    for name in ("learnAttentionGSOBatch", "learnAttentionGSOBatch_KeyQuery", "graphAttentionLSIGFBatch_KeyQuery",
                 "graphAttentionLSIGFBatch_modified"):
        assert str(inspect.signature(getattr(ours, name))) == str(inspect.signature(getattr(gml, name))), name
    ref_init = inspect.signature(gml.GraphFilterBatchAttentional.__init__)
    our_init = inspect.signature(ours.GraphFilterBatchAttentional.__init__)
    assert list(ref_init.parameters) == list(our_init.parameters)
    for k, v in ref_init.parameters.items():
        if k != "nonlinearity":
            assert our_init.parameters[k].default == v.default, k


def test_parameters_are_validated_before_the_kernels_see_them():
    """A layer left on the CPU, or cast to another dtype, must raise instead of handing raw pointers to CUDA."""
    from magat_pathplanning_b200 import graphML
    m = graphML.GraphFilterBatchAttentional(16, 16, 2, 2, 1, True, concatenate=True, attentionMode="KeyQuery")
    x = torch.zeros(1, 16, 5)
    for bad in (m.double(),):
        with pytest.raises(RuntimeError):
            graphML._check_params(x.float(), bad.filterWeight, bad.mixer, bad.weight, bad.weight_bias, bad.bias,
                                  need_cuda=False)
    m = m.float()
    graphML._check_params(x, m.filterWeight, m.mixer, m.weight, m.weight_bias, m.bias, need_cuda=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        graphML._check_params(x, m.filterWeight, m.mixer, m.weight, m.weight_bias, m.bias)


def test_module_surface_matches_reference():
    from magat_pathplanning_b200 import GraphFilterBatchAttentional
    m = GraphFilterBatchAttentional(16, 16, 3, 4, 1, True, concatenate=True, attentionMode="KeyQuery")
    assert [k for k, _ in m.named_parameters()] == ["mixer", "weight_bias", "filterWeight", "bias", "weight"]
    assert tuple(m.weight.shape) == (4, 1, 16, 16) and tuple(m.filterWeight.shape) == (4, 16, 1, 3, 16)
    assert float(m.weight_bias.abs().max()) == 0.0
    assert float(m.weight.abs().max()) <= 1.0 / (16 * 4) ** 0.5
    m2 = GraphFilterBatchAttentional(16, 24, 2, 2, 1, False, attentionMode="GAT_modified")
    assert "bias" not in m2.state_dict() and tuple(m2.weight.shape) == (2, 1, 24, 16)
    assert "no GSO stored" in repr(m2)
    with pytest.raises(AssertionError):
        m.addGSO(torch.zeros(2, 5, 5))
    m.addGSO(torch.zeros(2, 1, 5, 5))
    assert m.N == 5 and "number_nodes=5" in repr(m)
    assert m.aij is None
    with pytest.raises(AttributeError):
        GraphFilterBatchAttentional(8, 8, 2, 2, attentionMode="nonsense")


def test_reference_state_dict_loads():
    """Parameter names / shapes / order are API (checkpoints 'GFL.0.*'); load a reference layer's state."""
    from oracle.ref_loader import load_reference_graphml, reference_available
    if not reference_available():
        pytest.skip("reference checkout not present")
    gml = load_reference_graphml()
    from magat_pathplanning_b200 import GraphFilterBatchAttentional
    for mode in ("KeyQuery", "GAT_modified"):
        ref = gml.GraphFilterBatchAttentional(16, 16, 3, 2, 1, True, concatenate=True, attentionMode=mode)
        ours = GraphFilterBatchAttentional(16, 16, 3, 2, 1, True, concatenate=True, attentionMode=mode)
        assert list(ref.state_dict()) == list(ours.state_dict())
        ours.load_state_dict(ref.state_dict())
        assert repr(ours) == repr(ref)


def test_host_mask_packer_matches_numpy(built):
    """magat_gso_pack_host needs no GPU: |s| > 1e-9 per entry (NaN is no edge), bit j % 32 of word j / 32."""
    import numpy as np
    from magat_pathplanning_b200 import pack_gso_host
    gen = torch.Generator().manual_seed(3)
    for N, dtype in ((70, torch.float32), (64, torch.float64), (5, torch.float32)):
        S = torch.randn(3, 1, N, N, generator=gen).to(dtype) * (torch.rand(3, 1, N, N, generator=gen) < 0.2)
        S[0, 0, 0, 1] = float("nan")
        S[1, 0, 1, 0] = 5e-10
        S[2, 0, 2, 3] = -2e-9
        bits = pack_gso_host(S, threads=3).numpy().view(np.uint32)
        W = (N + 31) // 32
        ref = np.zeros((3 * N, W * 32), dtype=bool)
        ref[:, :N] = (S.abs() > 1e-9)[:, 0].reshape(3 * N, N).numpy()
        want = np.packbits(ref, axis=1, bitorder="little").view(np.uint32).reshape(3, N, W)
        assert np.array_equal(bits, want)
