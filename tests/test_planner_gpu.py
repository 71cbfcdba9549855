"""The reference's own CNN -> GAT -> MLP planner, unmodified, with the layer class rebound to the CUDA one
(INTEGRATION.md): logits against the same planner running the reference layer on the CPU, and the pickling round trip
the reference's multi-process simulation relies on (agents/decentralplannerlocal_OnlineExpert_GAT.py:720-728).

The reference files come from /root/reference or from baseline/_ref/ (baseline/fetch_reference.py; travels to the GPU
box).  Tolerance 1e-4 max-norm relative on the action logits (BASELINE.json north_star)."""
import io
import pickle

import pytest
import torch

from oracle.ref_loader import PlannerConfig, load_reference_planner, reference_available

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_available(), reason="no reference files (baseline/_ref)")]

TOL = 1e-4


@pytest.fixture(autouse=True)
def fp32_everywhere():
    """The planner's CNN and MLPs run in torch: keep cuDNN / cuBLAS off TF32 so the only difference between the two
    planners is the graph-attention layer under test."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def build_pair(model_file, cls, cfg_kw, seed=0):
    import magat_pathplanning_b200 as b200
    mod, gml = load_reference_planner(model_file)
    cfg_cpu = PlannerConfig.default(device="cpu", **cfg_kw)
    torch.manual_seed(seed)
    ref_model = getattr(mod, cls)(cfg_cpu).eval()
    originals = b200.install_into_reference(gml)
    try:
        cfg_gpu = PlannerConfig.default(device=torch.device("cuda:0"), **cfg_kw)
        our_model = getattr(mod, cls)(cfg_gpu)
    finally:
        for k, v in originals.items():
            setattr(gml, k, v)
    our_model.load_state_dict(ref_model.state_dict())
    return ref_model, our_model.to("cuda:0").eval()


def inputs(B, N, fov, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 3, fov + 2, fov + 2, generator=gen)
    pos = torch.rand(B, N, 2, generator=gen) * 20
    d = torch.cdist(pos, pos)
    A = ((d < 7.0) & ~torch.eye(N, dtype=torch.bool)).float()
    S = A / A.sum(-1).amax(-1).clamp_min(1.0)[:, None, None]
    return x, S


@pytest.mark.parametrize("model_file,cfg_kw,B,N", [
    ("decentralplanner_GAT", dict(attentionMode="KeyQuery", nGraphFilterTaps=3, nAttentionHeads=4, AttentionConcat=True), 3, 10),
    ("decentralplanner_GAT", dict(attentionMode="KeyQuery", nGraphFilterTaps=2, nAttentionHeads=1, AttentionConcat=False), 1, 10),
    ("decentralplanner_GAT", dict(attentionMode="GAT_modified", nGraphFilterTaps=3, nAttentionHeads=4, AttentionConcat=True), 2, 24),
    ("decentralplanner_GAT", dict(attentionMode="GAT_origin", nGraphFilterTaps=3, nAttentionHeads=4, AttentionConcat=True), 2, 12),
    ("decentralplanner_GAT_bottleneck_SkipConcat", dict(attentionMode="KeyQuery", nGraphFilterTaps=2, nAttentionHeads=4,
                                                        AttentionConcat=False, bottleneckFeature=32), 2, 16),
])
def test_reference_planner_forward_and_backward_through_the_cuda_layer(model_file, cfg_kw, B, N):
    ref_model, our_model = build_pair(model_file, "DecentralPlannerGATNet", cfg_kw)
    x, S = inputs(B, N, 9, seed=B * 100 + N)
    ref_model.addGSO(S.clone())
    ref_logits = ref_model(x)
    ref_logits.square().sum().backward()
    our_model.train(False)
    our_model.addGSO(S.clone().to("cuda:0"))
    our_logits = our_model(x.to("cuda:0"))
    assert our_logits.shape == ref_logits.shape == (B * N, 5)
    assert rel_err(our_logits, ref_logits) < TOL
    our_logits.square().sum().backward()
    gmax = max(float(p_.grad.abs().max()) for p_ in ref_model.parameters() if p_.grad is not None)
    for (k, pr), (_, po) in zip(ref_model.named_parameters(), our_model.named_parameters()):
        if pr.grad is None:
            assert po.grad is None, k
        elif float(pr.grad.abs().max()) < 1e-6 * gmax:
            # mathematically zero (GAT_modified's weight_bias shifts every score of a softmax row alike): rounding
            # noise in the reference too -- ours must be noise of the same size, not a value
            assert po.grad is not None and float(po.grad.abs().max()) < 1e-5 * gmax, k
        else:
            assert po.grad is not None and rel_err(po.grad, pr.grad) < 2e-4, k
    # inference under no_grad (the simulator loop, agents/...GAT.py:888-892) takes the single-launch small-graph kernel
    with torch.no_grad():
        our_model.addGSO(S.clone().to("cuda:0"))
        assert rel_err(our_model(x.to("cuda:0")), ref_logits) < TOL
    # the planner's tail -- GFL, actionsMLP, argmax decode -- as ONE call on the features the planner
    # feeds its graph layer (the fused head where the configuration allows it, layer + MLP otherwise)
    if len(our_model.GFL) == 1 and "Skip" not in model_file and hasattr(our_model.GFL[0], "forward_actions"):      # one graph layer (activation inside), no skip connection
        seen = {}
        hook = our_model.GFL[0].register_forward_pre_hook(lambda m, a: seen.__setitem__("x", a[0].detach()))
        with torch.no_grad():
            our_model.addGSO(S.clone().to("cuda:0"))
            our_model(x.to("cuda:0"))
            hook.remove()
            logits, keys = our_model.GFL[0].forward_actions(seen["x"], our_model.actionsMLP, return_actions=True)
        assert rel_err(logits, ref_logits) < TOL
        assert torch.equal(keys.long().cpu(), torch.max(torch.softmax(logits.cpu(), 1), 1)[1])


def _child(blob, x_np, S_np, q):
    try:
        import torch as t
        t.backends.cudnn.allow_tf32 = False            # as in the parent: the CNN / MLPs of the planner stay fp32
        t.backends.cuda.matmul.allow_tf32 = False
        from oracle.ref_loader import load_reference_planner as load
        load("decentralplanner_GAT")                   # registers graphs.models.* so the pickle resolves
        model = pickle.load(io.BytesIO(blob)).to("cuda:0").eval()
        with t.no_grad():
            model.addGSO(t.from_numpy(S_np).to("cuda:0"))
            q.put(("ok", model(t.from_numpy(x_np).to("cuda:0")).cpu().numpy()))
    except Exception as exc:                           # surface the failure in the parent instead of a bare EOFError
        import traceback
        q.put(("error", traceback.format_exc() + repr(exc)))


def test_planner_pickles_into_a_spawned_process():
    """The reference hands the model to spawned simulation workers; the rebound layer must survive that (no device
    scratch, ctypes handle or stream stored on the module)."""
    import torch.multiprocessing as mp
    cfg_kw = dict(attentionMode="KeyQuery", nGraphFilterTaps=3, nAttentionHeads=4, AttentionConcat=True)
    ref_model, our_model = build_pair("decentralplanner_GAT", "DecentralPlannerGATNet", cfg_kw)
    x, S = inputs(2, 10, 9, seed=5)
    ref_model.addGSO(S.clone())
    with torch.no_grad():
        ref_logits = ref_model(x)
        our_model.addGSO(S.clone().to("cuda:0"))
        our_model(x.to("cuda:0"))                      # leaves device-side state on the layer before pickling
    our_model.S = None                                 # (the reference pickles the model between episodes)
    buf = io.BytesIO()
    pickle.dump(our_model.cpu(), buf)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pr = ctx.Process(target=_child, args=(buf.getvalue(), x.numpy(), S.numpy(), q))
    pr.start()
    status, out = q.get(timeout=180)
    pr.join(timeout=60)
    assert status == "ok", out
    assert pr.exitcode == 0
    assert rel_err(torch.from_numpy(out), ref_logits) < TOL
