"""The unmodified reference planner builds on top of the rebound layer class (CPU, reference checkout present)."""
import pytest
import torch

from oracle.ref_loader import PlannerConfig, load_reference_planner, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference checkout not present")


@pytest.mark.parametrize("model_file,cls", [("decentralplanner_GAT", "DecentralPlannerGATNet"),
                                            ("decentralplanner_GAT_bottleneck_SkipConcat", "DecentralPlannerGATNet")])
def test_reference_planner_builds_with_our_layer(model_file, cls):
    import magat_pathplanning_b200 as b200
    mod, gml = load_reference_planner(model_file)
    cfg = PlannerConfig.default()
    torch.manual_seed(0)
    ref_model = getattr(mod, cls)(cfg)
    originals = b200.install_into_reference(gml)
    try:
        torch.manual_seed(0)
        our_model = getattr(mod, cls)(cfg)
    finally:
        for k, v in originals.items():
            setattr(gml, k, v)
    assert type(our_model.GFL[0]).__module__.startswith("magat_pathplanning_b200")
    ref_sd, our_sd = ref_model.state_dict(), our_model.state_dict()
    assert list(ref_sd) == list(our_sd)
    assert all(ref_sd[k].shape == our_sd[k].shape for k in ref_sd)
    our_model.load_state_dict(ref_sd)                      # reference checkpoints load unchanged
    assert repr(our_model.GFL[0]) == repr(ref_model.GFL[0])
    # model-level addGSO reaches the layer untouched (decentralplanner_GAT.py:260-276, :321)
    S = torch.rand(2, 10, 10)
    our_model.addGSO(S)
    assert our_model.S.shape == (2, 1, 10, 10)
