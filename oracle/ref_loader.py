"""Load the UNMODIFIED reference ``graphML.py`` (and planner model files) by file path.  TEST INFRASTRUCTURE ONLY.

Looks for the reference in ``$MAGAT_REFERENCE``, then ``/root/reference`` (the build container's read-only checkout),
then ``baseline/_ref/`` -- the git-ignored copy of the few files the path needs that ``baseline/fetch_reference.py``
makes and that travels to the GPU box with the snapshot.  Used by ``tests/golden/make_golden.py`` (golden vectors), by
the tests that compare the oracle / the CUDA layer with the live reference, and by ``bench.py --impl reference``.
Nothing under ``magat_pathplanning_b200/`` imports it.

``utils/__init__.py`` of the reference auto-imports every module (matplotlib, easydict,
... are absent here), so the packages are stubbed and the one file is exec'd directly
(recipe: SURVEY.md section 8c).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    for cand in (os.environ.get("MAGAT_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "utils", "graphUtils", "graphML.py")):
            return cand
    return "/root/reference"


DEFAULT_REF = _find_reference()


def reference_available(ref: str = DEFAULT_REF) -> bool:
    return os.path.isfile(os.path.join(ref, "utils", "graphUtils", "graphML.py"))


def load_reference_graphml(ref: str = DEFAULT_REF):
    name = "utils.graphUtils.graphML"
    if name in sys.modules and getattr(sys.modules[name], "__magat_ref__", False):
        return sys.modules[name]
    for pkg in ("utils", "utils.graphUtils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    gt = types.ModuleType("utils.graphUtils.graphTools")   # imported at graphML.py:43, unused on this path
    sys.modules.setdefault("utils.graphUtils.graphTools", gt)
    sys.modules["utils.graphUtils"].graphTools = sys.modules["utils.graphUtils.graphTools"]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(ref, "utils", "graphUtils", "graphML.py"))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # '\s' in reference docstrings
        spec.loader.exec_module(mod)
    mod.__magat_ref__ = True
    sys.modules[name] = mod
    sys.modules["utils.graphUtils"].graphML = mod
    return mod


def load_reference_planner(name: str = "decentralplanner_GAT", ref: str = DEFAULT_REF):
    """Load one of the reference's planner model files (graphs/models/<name>.py) by path with the
    packages it imports stubbed (recipe: SURVEY.md section 8c).  Build-container only."""
    gml = load_reference_graphml(ref)
    for pkg in ("graphs", "graphs.models"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    if "torchsummaryX" not in sys.modules:
        ts = types.ModuleType("torchsummaryX")
        ts.summary = lambda *a, **k: None          # imported at decentralplanner_GAT.py:11, never called
        sys.modules["torchsummaryX"] = ts

    def by_path(modname, relpath):
        if modname in sys.modules and getattr(sys.modules[modname], "__magat_ref__", False):
            return sys.modules[modname]
        spec = importlib.util.spec_from_file_location(modname, os.path.join(ref, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
        mod.__magat_ref__ = True
        return mod

    by_path("graphs.weights_initializer", "graphs/weights_initializer.py")
    by_path("graphs.models.resnet_pytorch", "graphs/models/resnet_pytorch.py")
    return by_path("graphs.models." + name, f"graphs/models/{name}.py"), gml


class PlannerConfig(dict):
    """Attribute dict with the fields the planners read (graphs/models/decentralplanner_GAT.py)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    @classmethod
    def default(cls, **kw):
        c = cls(num_agents=10, map_w=20, map_h=20, FOV=9, numInputFeatures=128, nGraphFilterTaps=3,
                nAttentionHeads=4, use_dropout=False, CNN_mode="Default", attentionMode="KeyQuery",
                AttentionConcat=True, GSO_mode="dist_GSO", device="cpu", bottleneckFeature=32,
                bottleneckMode=None, batch_numAgent=False, return_attentionGSO=False)
        c.update(kw)
        return c
