"""Load the UNMODIFIED reference ``graphML.py`` by file path.  TEST INFRASTRUCTURE ONLY.

Only usable where the reference checkout exists (the build container, /root/reference);
it does not exist on the GPU box, so nothing in the ``-m gpu`` tests, ``smoke()`` or
``bench.py`` calls this.  It is used by ``tests/golden/make_golden.py`` to produce the
committed golden vectors and by the optional CPU tests that compare the oracle against
the live reference when it is present.

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

DEFAULT_REF = os.environ.get("MAGAT_REFERENCE", "/root/reference")


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
