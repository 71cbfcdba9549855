"""B200-native (sm_100a) batched graph-attention layer of proroklab/magat_pathplanning.

Only the hot path is here: ``GraphFilterBatchAttentional`` and the functionals it calls
(reference: utils/graphUtils/graphML.py:4506-4685, :1724-1827, :1180-1286, :713-823), behind the
reference's own module interface, plus the non-attentional ``GraphFilterBatch`` / ``BatchLSIGF`` (:5485-5700) on
the same kernels.  See DESIGN.md / INTEGRATION.md.
"""
from .graphML import (GraphFilterBatchAttentional, graphAttentionLSIGFBatch_KeyQuery,  # noqa: F401
                      graphAttentionLSIGFBatch_modified, learnAttentionGSOBatch_KeyQuery,
                      learnAttentionGSOBatch, build_adjacency, build_adjacency_from_positions, gat_layer, gat_layer_actions,
                      attention_dense, GraphFilterBatch, BatchLSIGF, pack_gso_host, build_adjacency_host,
                      build_adjacency_from_rowbits, GraphFilterBatchAttentional_Origin,
                      graphAttentionLSIGFBatch_Origin, learnAttentionGSOBatch_origin)
from .integration import install_into_reference  # noqa: F401

__all__ = ["GraphFilterBatchAttentional", "graphAttentionLSIGFBatch_KeyQuery",
           "graphAttentionLSIGFBatch_modified", "learnAttentionGSOBatch_KeyQuery", "learnAttentionGSOBatch",
           "build_adjacency", "build_adjacency_from_positions", "gat_layer", "gat_layer_actions", "attention_dense", "install_into_reference",
           "GraphFilterBatch", "BatchLSIGF", "pack_gso_host", "build_adjacency_host", "build_adjacency_from_rowbits",
           "GraphFilterBatchAttentional_Origin", "graphAttentionLSIGFBatch_Origin", "learnAttentionGSOBatch_origin"]
