"""Swap the B200 layer into an imported copy of the reference.

The planners construct the layer as ``gml.GraphFilterBatchAttentional(...)`` with
``import utils.graphUtils.graphML as gml`` (graphs/models/decentralplanner_GAT.py:191-192 and
its seven siblings), so rebinding that one module attribute is the whole integration.
"""
from __future__ import annotations

import sys


def install_into_reference(graphml_module=None):
    """Rebind ``GraphFilterBatchAttentional`` (and the four functionals on its path) inside the
    reference's ``utils.graphUtils.graphML`` module.  Returns the originals for restoring."""
    from . import graphML as ours
    mod = graphml_module if graphml_module is not None else sys.modules.get("utils.graphUtils.graphML")
    if mod is None:
        raise RuntimeError("import utils.graphUtils.graphML (the reference) before installing")
    names = ("GraphFilterBatchAttentional", "graphAttentionLSIGFBatch_KeyQuery",
             "graphAttentionLSIGFBatch_modified", "learnAttentionGSOBatch_KeyQuery", "learnAttentionGSOBatch",
             "GraphFilterBatch", "BatchLSIGF",
             "GraphFilterBatchAttentional_Origin", "graphAttentionLSIGFBatch_Origin", "learnAttentionGSOBatch_origin")
    originals = {n: getattr(mod, n, None) for n in names}
    for n in names:
        setattr(mod, n, getattr(ours, n))
    return originals
