"""Install the native path into an importable copy of the reference, so its ``train.py`` /
``infer.py`` run unmodified.  Usage, before the reference builds its model::

    import hpcs_b200.patch; hpcs_b200.patch.install()

Only module attributes are rebound; nothing of the reference is edited on disk.
"""
from __future__ import annotations

import importlib
import sys

from . import decode, graph, hyperbolic, loss

# (reference module, attribute) -> replacement
TARGETS = {
    ("hpcs.nn.dgcnn.utils.vn_dgcnn_util", "knn"): graph.knn,
    ("hpcs.nn.dgcnn.utils.vn_dgcnn_util", "get_graph_feature"): graph.get_graph_feature,
    ("hpcs.nn.dgcnn.utils.vn_dgcnn_util", "get_graph_feature_cross"): graph.get_graph_feature_cross,
    ("hpcs.nn.dgcnn.vn_dgcnn_partseg", "get_graph_feature"): graph.get_graph_feature,
    ("hpcs.nn.pointnet.utils.vn_dgcnn_util", "knn"): graph.knn,
    ("hpcs.nn.pointnet.utils.vn_dgcnn_util", "get_graph_feature_cross"): graph.get_graph_feature_cross,
    ("hpcs.distances.lca", "hyp_lca"): hyperbolic.hyp_lca,
    ("hpcs.distances", "hyp_lca"): hyperbolic.hyp_lca,
    ("hpcs.nn.hyperbolic.hyp_embed", "ExpMap"): hyperbolic.ExpMap,
    ("hpcs.nn.hyperbolic", "ExpMap"): hyperbolic.ExpMap,
    ("hpcs.miner.loss_and_miner_utils", "get_balanced_random_triplet_indices"): loss.get_balanced_random_triplet_indices,
    ("hpcs.miner.triplet_margin_miner", "RandomTripletMarginMiner"): loss.RandomTripletMarginMiner,
    ("hpcs.loss.ultrametric_loss", "MetricHyperbolicLoss"): loss.MetricHyperbolicLoss,
    ("hpcs.loss", "MetricHyperbolicLoss"): loss.MetricHyperbolicLoss,
    ("hpcs.models.base_hyp_hc", "MetricHyperbolicLoss"): loss.MetricHyperbolicLoss,
    ("hpcs.utils.scores", "get_optimal_k"): decode.get_optimal_k,
    ("hpcs.models.base_hyp_hc", "get_optimal_k"): decode.get_optimal_k,
}


def _decode_linkage(self, leaves_embeddings):
    """Bound onto ``BaseSimilarityHypHC`` (hpcs/models/base_hyp_hc.py:81-86)."""
    return decode.decode_linkage(leaves_embeddings, self.metric_hyp_loss.scale, method="complete")


def install(strict: bool = False) -> list:
    """Rebind every target whose module can be imported; returns the list of patched names."""
    done = []
    for (mod_name, attr), repl in TARGETS.items():
        try:
            mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        setattr(mod, attr, repl)
        done.append(f"{mod_name}.{attr}")
    try:
        base = sys.modules.get("hpcs.models.base_hyp_hc") or importlib.import_module("hpcs.models.base_hyp_hc")
        base.BaseSimilarityHypHC._decode_linkage = _decode_linkage
        done.append("hpcs.models.base_hyp_hc.BaseSimilarityHypHC._decode_linkage")
    except Exception:
        if strict:
            raise
    return done
