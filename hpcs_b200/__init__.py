"""hpcs_b200 -- B200 (sm_100a) native hot path for HPCS behind the reference's Python signatures.

kNN graph + edge features, Poincare-ball triplet objective, linkage decode.  CUDA only: importing
is cheap and works anywhere, but every op raises unless its tensors live on a CC 10.x device.
"""
from .graph import knn, get_graph_feature, get_graph_feature_cross
from .hyperbolic import hyp_lca, expmap0, ExpMap, normalize_project
from .loss import (CosineSimilarity, RandomTripletMarginMiner, MetricHyperbolicLoss, CosFaceLoss,
                   HierarchicalCosFaceLoss, HierarchicalMetricHyperbolicLoss, hierarchical_loss,
                   get_balanced_random_triplet_indices, hyp_triplet_loss, filter_triplets,
                   sample_triplets_device, sampler_state, triplet_segments, triplet_plan)
from .pipeline import rotate_points, rotation_params, to_categorical
from .decode import decode_linkage, decode_linkage_batch, linkage_from_leaves, fcluster_maxclust, get_optimal_k_batch, get_optimal_k

__all__ = [
    "knn", "get_graph_feature", "get_graph_feature_cross",
    "hyp_lca", "expmap0", "ExpMap", "normalize_project",
    "CosineSimilarity", "RandomTripletMarginMiner", "MetricHyperbolicLoss", "CosFaceLoss",
    "HierarchicalCosFaceLoss", "HierarchicalMetricHyperbolicLoss", "hierarchical_loss",
    "rotate_points", "rotation_params", "to_categorical",
    "get_balanced_random_triplet_indices", "hyp_triplet_loss", "filter_triplets",
    "sample_triplets_device", "sampler_state", "triplet_segments", "triplet_plan",
    "decode_linkage", "decode_linkage_batch", "linkage_from_leaves", "fcluster_maxclust", "get_optimal_k_batch", "get_optimal_k",
]
