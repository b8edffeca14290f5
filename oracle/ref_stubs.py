"""Import shims that let the UNMODIFIED reference (``/root/reference``) run in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (and nothing else) to execute the
reference's own code on CPU and record golden vectors under ``tests/golden/``.

The reference depends on ``pytorch-metric-learning`` 1.6.3 (``hpcs-env.yaml:283``), which is not
installed here and cannot be (no network).  Only a handful of its base classes are touched on the
hot path; their published behaviour is restated below (nothing is copied from the reference):

* ``BaseDistance.__call__``: L2-normalise rows (``F.normalize``, p=2, eps 1e-12) when
  ``normalize_embeddings`` is set, call ``compute_mat``, raise to ``power`` if != 1.
  Call sites: ``hpcs/distances/cosine.py:4-16``, ``hpcs/miner/triplet_margin_miner.py:16``,
  ``hpcs/loss/ultrametric_loss.py:65``.
* ``DotProductSimilarity``: ``is_inverted=True``; ``compute_mat = q @ r.T``.
* ``BaseMiner.forward``: run ``mine`` under ``no_grad`` with ``ref_emb = embeddings`` when no
  reference set is given.  Call site: ``hpcs/miner/triplet_margin_miner.py:6-13``.
* ``BaseMetricLossFunction`` / ``CosFaceLoss`` / ``LargeMarginSoftmaxLoss`` / reducers: only need
  to be constructible so that ``hpcs.loss.ultrametric_loss`` imports; the CosFace term is outside
  the hot path and is not used for goldens.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def _module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


class _BaseDistance(torch.nn.Module):
    def __init__(self, normalize_embeddings=True, p=2, power=1, is_inverted=False, **kwargs):
        super().__init__()
        self.normalize_embeddings = normalize_embeddings
        self.p = p
        self.power = power
        self.is_inverted = is_inverted

    def forward(self, query_emb, ref_emb=None):
        q = self.maybe_normalize(query_emb)
        r = q if ref_emb is None else self.maybe_normalize(ref_emb)
        mat = self.compute_mat(q, r)
        if self.power != 1:
            mat = mat ** self.power
        return mat

    def normalize(self, embeddings, dim=1, **kwargs):
        return F.normalize(embeddings, p=self.p, dim=dim, **kwargs)

    def maybe_normalize(self, embeddings, dim=1, **kwargs):
        if self.normalize_embeddings:
            return self.normalize(embeddings, dim=dim, **kwargs)
        return embeddings

    def smallest_dist(self, *args, **kwargs):
        return torch.max(*args, **kwargs) if self.is_inverted else torch.min(*args, **kwargs)

    def margin(self, x, y):
        return y - x if self.is_inverted else x - y


class _DotProductSimilarity(_BaseDistance):
    def __init__(self, **kwargs):
        super().__init__(is_inverted=True, **kwargs)

    def compute_mat(self, query_emb, ref_emb):
        return torch.matmul(query_emb, ref_emb.t())

    def pairwise_distance(self, query_emb, ref_emb):
        return torch.sum(query_emb * ref_emb, dim=1)


class _BaseMiner(torch.nn.Module):
    def __init__(self, distance=None, **kwargs):
        super().__init__()
        self.distance = distance

    def forward(self, embeddings, labels, ref_emb=None, ref_labels=None):
        with torch.no_grad():
            labels = labels.to(embeddings.device)
            if ref_emb is None:
                ref_emb, ref_labels = embeddings, labels
            return self.mine(embeddings, labels, ref_emb, ref_labels)


class _TripletMarginMiner(_BaseMiner):
    def __init__(self, margin=0.2, type_of_triplets="all", **kwargs):
        super().__init__(**kwargs)
        self.margin = margin
        self.type_of_triplets = type_of_triplets


class _BaseMetricLossFunction(torch.nn.Module):
    def __init__(self, distance=None, **kwargs):
        super().__init__()
        self.distance = distance

    def add_to_recordable_attributes(self, *args, **kwargs):
        pass


class _LargeMarginSoftmaxLoss(_BaseMetricLossFunction):
    def __init__(self, num_classes=2, embedding_size=2, margin=4, scale=1, **kwargs):
        super().__init__(**kwargs)
        self.margin, self.scale = margin, scale
        self.W = torch.nn.Parameter(torch.randn(embedding_size, num_classes))


class _Reducer(torch.nn.Module):
    pass


def install() -> None:
    """Register the shim modules and put the reference on ``sys.path`` (idempotent)."""
    if "pytorch_metric_learning" in sys.modules:
        return
    pml = _module("pytorch_metric_learning")
    dist = _module("pytorch_metric_learning.distances")
    dist.BaseDistance = _BaseDistance
    dist.DotProductSimilarity = _DotProductSimilarity
    miners = _module("pytorch_metric_learning.miners")
    miners.BaseMiner = _BaseMiner
    miners.TripletMarginMiner = _TripletMarginMiner
    losses = _module("pytorch_metric_learning.losses")
    losses.BaseMetricLossFunction = _BaseMetricLossFunction
    losses.TripletMarginLoss = _BaseMetricLossFunction
    losses.LargeMarginSoftmaxLoss = _LargeMarginSoftmaxLoss
    losses.CosFaceLoss = _LargeMarginSoftmaxLoss
    base = _module("pytorch_metric_learning.losses.base_metric_loss_function")
    base.BaseMetricLossFunction = _BaseMetricLossFunction
    reducers = _module("pytorch_metric_learning.reducers")
    reducers.AvgNonZeroReducer = _Reducer
    utils = _module("pytorch_metric_learning.utils")
    c_f = _module("pytorch_metric_learning.utils.common_functions")
    c_f.to_device = lambda x, tensor=None, device=None, dtype=None: x.to(
        device=device if device is not None else tensor.device, dtype=dtype if dtype is not None else x.dtype)
    lmu = _module("pytorch_metric_learning.utils.loss_and_miner_utils")
    utils.common_functions = c_f
    utils.loss_and_miner_utils = lmu
    pml.distances, pml.miners, pml.losses, pml.reducers, pml.utils = dist, miners, losses, reducers, utils
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_by_path(name: str, relpath: str):
    """Import one reference file by path (skips package ``__init__`` files that need geoopt)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, f"{REFERENCE_ROOT}/{relpath}")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
